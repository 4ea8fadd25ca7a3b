// Kernels 1 + 2 of the north star: sort-based broadphase and narrowphase, plus Contact.set.
//
// Replaces CollisionProcessor.collisionDetection (CollisionProcessor.java:91-102): the O(N^2) pair loop
// broadPhase :683-697 becomes a Morton-keyed uniform grid (radix sort of cell codes + neighbour-cell
// sweep over bounding spheres); narrowPhase :707-757 becomes one kernel per primitive-pair class writing
// into per-pair slots that are compacted in canonical (bodyLo, bodyHi, partLo, partHi, emission) order.
// The broadphase is a CONSERVATIVE filter: it only drops pairs that the reference's own primitives
// reject on their first test (BoxBox.java:466, BoxSphere.java:157, BoxPlane.java:29,
// CollisionProcessor.java:801,985), so the emitted contact set is identical.
#pragma once
#include <cub/cub.cuh>

#include "am3d_collide.cuh"
#include "am3d_ctx.h"

// ------------------------------------------------------------------------------------------------
// shape world transforms and bounding spheres
// ------------------------------------------------------------------------------------------------
__global__ void k_shape_update(int nsh, const int* __restrict__ shType, const int* __restrict__ shBody,
                               const int* __restrict__ shRoot, const double* __restrict__ shRadius,
                               const double* __restrict__ shSize, const double* __restrict__ shLR, const double* __restrict__ shLt,
                               const int* __restrict__ btype, const double* __restrict__ x, const double* __restrict__ R,
                               const double* __restrict__ ndC, const double* __restrict__ ndR, double* __restrict__ shX,
                               double* __restrict__ shR, double* __restrict__ shBoundC, double* __restrict__ shBoundR,
                               double* __restrict__ shBoundH) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsh) return;
  int b = shBody[s];
  xf T;
  T.R = ldm(R + 9 * b);
  T.t = ld3(x + 3 * b);
  if (btype[b] == AM3D_BODY_COMPOSITE) {
    // RigidBodyGeomComposite.updateBodyPositionsFromParent :32-36
    xf L;
    L.R = ldm(shLR + 9 * s);
    L.t = ld3(shLt + 3 * s);
    T = xfMul(T, L);
  }
  stm(shR + 9 * s, T.R);
  st3(shX + 3 * s, T.t);
  int t = shType[s];
  // bounding sphere (grid cell size, plane tests) and world AABB half extents (pair filter)
  if (t == AM3D_SHAPE_BOX) {
    st3(shBoundC + 3 * s, T.t);
    shBoundR[s] = shRadius[s];
    d3 h = vscale(0.5, ld3(shSize + 3 * s));
    const m3& A = T.R;
    st3(shBoundH + 3 * s, d3(fabs(A.m[0]) * h.x + fabs(A.m[1]) * h.y + fabs(A.m[2]) * h.z,
                             fabs(A.m[3]) * h.x + fabs(A.m[4]) * h.y + fabs(A.m[5]) * h.z,
                             fabs(A.m[6]) * h.x + fabs(A.m[7]) * h.y + fabs(A.m[8]) * h.z));
  } else if (t == AM3D_SHAPE_TREE) {
    int root = shRoot[s];
    st3(shBoundC + 3 * s, xfP(T, ld3(ndC + 3 * root)));
    shBoundR[s] = ndR[root];
    st3(shBoundH + 3 * s, d3(ndR[root], ndR[root], ndR[root]));
  } else {
    st3(shBoundC + 3 * s, d3());
    shBoundR[s] = 0;
    st3(shBoundH + 3 * s, d3());
  }
}

// ------------------------------------------------------------------------------------------------
// broadphase
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long spread21(unsigned int v) {
  unsigned long long x = v & 0x1fffff;
  x = (x | x << 32) & 0x1f00000000ffffULL;
  x = (x | x << 16) & 0x1f0000ff0000ffULL;
  x = (x | x << 8) & 0x100f00f00f00f00fULL;
  x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
  x = (x | x << 2) & 0x1249249249249249ULL;
  return x;
}
// 14 bits per axis Morton code + scene id in the high bits
__device__ __forceinline__ unsigned long long cellCode(int ix, int iy, int iz, int scene) {
  return ((unsigned long long)scene << 42) | spread21(ix) | (spread21(iy) << 1) | (spread21(iz) << 2);
}
__device__ __forceinline__ int cellCoord(double c, double inv) {
  double f = floor(c * inv) + 8192.0;
  f = fmin(fmax(f, 0.0), 16383.0);
  return (int)f;
}

__global__ void k_cell_keys(int n, const int* __restrict__ list, const int* __restrict__ shBody,
                            const int* __restrict__ scene, const double* __restrict__ bc, double inv,
                            unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = list[i];
  d3 c = ld3(bc + 3 * s);
  keys[i] = cellCode(cellCoord(c.x, inv), cellCoord(c.y, inv), cellCoord(c.z, inv), scene[shBody[s]]);
  vals[i] = s;
}

struct PairCtx {
  const int* shBody;
  const int* bShapeFirst;
  const int* parent;
  const int* flags;
  const int* scene;
  const long long* stamp;
  const double* bc;
  const double* br;
  const double* bh;
  unsigned long long* pairKey;
  unsigned long long* pairVal;
  int* counter;
  int cap;
  int bodyBits, partBits;  // the pair SORT key is (bodyLo, bodyHi[, partLo, partHi]) packed without gaps: fewer radix passes
};
// canonical contact-order key of a shape pair
__device__ __forceinline__ unsigned long long pairKey0(const int* __restrict__ shBody, const int* __restrict__ bShapeFirst, int sa, int sb) {
  int ba = shBody[sa], bb = shBody[sb];
  int slo = ba < bb ? sa : sb, shi = ba < bb ? sb : sa;
  int blo = shBody[slo], bhi = shBody[shi];
  unsigned long long plo = slo - bShapeFirst[blo], phi = shi - bShapeFirst[bhi];
  return ((unsigned long long)blo << 40) | ((unsigned long long)bhi << 16) | (plo << 8) | phi;
}

// top-level filter of broadPhase :692-693 + pair orientation (list order of RigidBodySystem.bodies)
__device__ __forceinline__ void tryPair(const PairCtx& C, int sa, int sb) {
  int ba = C.shBody[sa], bb = C.shBody[sb];
  if (ba == bb) return;
  if (C.scene[ba] != C.scene[bb]) return;
  int ta = C.parent[ba] >= 0 ? C.parent[ba] : ba;
  int tb = C.parent[bb] >= 0 ? C.parent[bb] : bb;
  if (ta == tb) return;
  int fa = C.flags[ta], fb = C.flags[tb];
  if ((fa | fb) & AM3D_F_DORMANT) return;  // not in RigidBodySystem.bodies
  bool pa = fa & AM3D_F_PINNED, pb = fb & AM3D_F_PINNED;
  if (pa && pb) return;
  if ((pa && (fb & AM3D_F_SLEEPING)) || (pb && (fa & AM3D_F_SLEEPING))) return;
  int idx = atomicAdd(C.counter, 1);
  if (idx >= C.cap) return;
  bool aFirst = C.stamp[ta] < C.stamp[tb];
  int first = aFirst ? sa : sb, second = aFirst ? sb : sa;
  int slo = ba < bb ? sa : sb, shi = ba < bb ? sb : sa;
  int blo = C.shBody[slo], bhi = C.shBody[shi];
  unsigned long long plo = slo - C.bShapeFirst[blo], phi = shi - C.bShapeFirst[bhi];
  unsigned long long k = ((unsigned long long)blo << C.bodyBits) | (unsigned long long)bhi;
  if (C.partBits) k = (k << 16) | (plo << 8) | phi;
  C.pairKey[idx] = k;
  C.pairVal[idx] = ((unsigned long long)(unsigned)first << 32) | (unsigned)second;
}

// World AABBs (boxes: |R| h, trees: the root sphere's cube) with a margin far above rounding error: boxes whose AABBs
// are apart are disjoint, so the reference's own tests (bounding spheres, then the 15-axis SAT; root-sphere tests for
// trees) return no contact for them -- the filter never changes a contact set.
__device__ __forceinline__ bool aabbOverlap(const d3& ca, const d3& ha, const d3& cb, const d3& hb) {
  const double rel = 1.0 + 1e-12, abs_ = 1e-9;
  return fabs(ca.x - cb.x) <= (ha.x + hb.x) * rel + abs_ && fabs(ca.y - cb.y) <= (ha.y + hb.y) * rel + abs_ &&
         fabs(ca.z - cb.z) <= (ha.z + hb.z) * rel + abs_;
}

// one thread per small shape: its own cell (partners with a larger shape id) and the 13 neighbouring cells of the
// "positive" half space (all partners), located in the sorted cell-code array: every pair is met exactly once
__global__ void k_pairs_grid(int n, const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                             double inv, PairCtx C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = vals[i];
  d3 c = ld3(C.bc + 3 * s), h = ld3(C.bh + 3 * s);
  int sc = C.scene[C.shBody[s]];
  int ix = cellCoord(c.x, inv), iy = cellCoord(c.y, inv), iz = cellCoord(c.z, inv);
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        const bool own = dx == 0 && dy == 0 && dz == 0;
        if (!own && (dz < 0 || (dz == 0 && (dy < 0 || (dy == 0 && dx < 0))))) continue;
        int jx = ix + dx, jy = iy + dy, jz = iz + dz;
        if (jx < 0 || jy < 0 || jz < 0 || jx > 16383 || jy > 16383 || jz > 16383) continue;
        unsigned long long k = cellCode(jx, jy, jz, sc);
        int lo = 0, hi = n;
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (keys[mid] < k) lo = mid + 1; else hi = mid;
        }
        for (int j = lo; j < n && keys[j] == k; j++) {
          int t = vals[j];
          if (own && t <= s) continue;
          if (!aabbOverlap(c, h, ld3(C.bc + 3 * t), ld3(C.bh + 3 * t))) continue;
          tryPair(C, s, t);
        }
      }
}

// one thread per non-plane shape: test against the (few) large shapes and the planes
__global__ void k_pairs_special(int nsh, const int* __restrict__ shType, const int* __restrict__ shLarge,
                                const int* __restrict__ largeStart, const int* __restrict__ largeList,
                                const int* __restrict__ planeStart, const int* __restrict__ planeList,
                                const double* __restrict__ shSize, const double* __restrict__ shRadius, PairCtx C) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsh) return;
  int ty = shType[t];
  if (ty == AM3D_SHAPE_PLANE) return;
  d3 c = ld3(C.bc + 3 * t), hh = ld3(C.bh + 3 * t);
  double r = C.br[t];
  int sc = C.scene[C.shBody[t]];
  for (int k = largeStart[sc]; k < largeStart[sc + 1]; k++) {
    int l = largeList[k];
    if (l == t) continue;
    if (shLarge[t] && t < l) continue;  // large-large pairs once
    if (!aabbOverlap(c, hh, ld3(C.bc + 3 * l), ld3(C.bh + 3 * l))) continue;
    tryPair(C, t, l);
  }
  for (int k = planeStart[sc]; k < planeStart[sc + 1]; k++) {
    int pl = planeList[k];
    d3 n = ld3(shSize + 3 * pl);
    double d = shRadius[pl];
    // same first test as BoxPlane.java:29 / CollisionProcessor.java:801
    if (ty == AM3D_SHAPE_BOX) {
      if (c.x * n.x + c.y * n.y + c.z * n.z + d > r) continue;
    } else {
      if (!(n.x * c.x + n.y * c.y + n.z * c.z + d - r < 0)) continue;
    }
    tryPair(C, t, pl);
  }
}

// ------------------------------------------------------------------------------------------------
// narrowphase
// ------------------------------------------------------------------------------------------------
enum { PT_BOXBOX = 0, PT_PLANEBOX = 1, PT_TREEPLANE = 2, PT_BOXTREE = 3, PT_TREETREE = 4 };

__global__ void k_pair_classify(int np, unsigned long long* __restrict__ pairVal, const int* __restrict__ shType,
                                int* __restrict__ pairType, int* __restrict__ pairCap, int* __restrict__ treeList,
                                int* __restrict__ treeCount) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  int a = (int)(pairVal[i] >> 32), b = (int)(pairVal[i] & 0xffffffffu);
  int ta = shType[a], tb = shType[b];
  int t;
  bool swap = false;
  // after this kernel (a,b) = (shape of Contact.body1, shape of Contact.body2): plane first against a box
  // (BoxPlane.java:50), tree first against a plane (CollisionProcessor.java:817), box first against a tree
  // (:742,748); box-box and tree-tree keep the list order of RigidBodySystem.bodies.
  if (ta == AM3D_SHAPE_PLANE || tb == AM3D_SHAPE_PLANE) {
    int o = ta == AM3D_SHAPE_PLANE ? tb : ta;
    if (o == AM3D_SHAPE_BOX) { t = PT_PLANEBOX; swap = ta != AM3D_SHAPE_PLANE; }
    else { t = PT_TREEPLANE; swap = ta == AM3D_SHAPE_PLANE; }
  } else if (ta == AM3D_SHAPE_BOX && tb == AM3D_SHAPE_BOX) {
    t = PT_BOXBOX;
  } else if (ta == AM3D_SHAPE_BOX || tb == AM3D_SHAPE_BOX) {
    t = PT_BOXTREE;
    swap = ta != AM3D_SHAPE_BOX;
  } else {
    t = PT_TREETREE;
  }
  if (swap) pairVal[i] = ((unsigned long long)(unsigned)b << 32) | (unsigned)a;
  pairType[i] = t;
  pairCap[i] = (t == PT_BOXBOX || t == PT_PLANEBOX) ? 8 : 0;  // tree pairs: filled by the count pass
  if (t != PT_BOXBOX && t != PT_PLANEBOX) treeList[atomicAdd(treeCount, 1)] = i;  // order is irrelevant: every pair owns its slots
}

struct HitOut {
  double* pos;
  double* nrm;
  double* viol;
  int* meta;
};
__device__ __forceinline__ void writeHit(const HitOut& H, long long slot, const d3& p, const d3& n, double viol, int info,
                                         int bv1, int bv2, int leaf) {
  st3(H.pos + 3 * slot, p);
  st3(H.nrm + 3 * slot, n);
  H.viol[slot] = viol;
  int4 m = make_int4(info, bv1, bv2, leaf);
  reinterpret_cast<int4*>(H.meta)[slot] = m;
}

struct HitSlotSink {
  const HitOut& H;
  long long base;
  __device__ __forceinline__ void set(int k, const d3& p, const d3& n, int info, double viol) const {
    writeHit(H, base + k, p, n, viol, info, AM3D_BV_NULL, AM3D_BV_NULL, -1);
  }
};

// box x box and plane x box: one thread per pair
__global__ void k_narrow_box(int np, const unsigned long long* __restrict__ pairVal, const int* __restrict__ pairType,
                             const int* __restrict__ pairSlot, const double* __restrict__ shSize, const double* __restrict__ shRadius,
                             const double* __restrict__ shX, const double* __restrict__ shR, HitOut H,
                             int* __restrict__ pairCount) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  int pt = pairType[i];
  if (pt != PT_BOXBOX && pt != PT_PLANEBOX) return;
  int a = (int)(pairVal[i] >> 32), b = (int)(pairVal[i] & 0xffffffffu);
  HitSlotSink sink{H, pairSlot[i]};  // hits go straight to the pair's slots: no per-thread staging array in local memory
  int n;
  if (pt == PT_BOXBOX) {
    n = collideBoxBox(ld3(shX + 3 * a), ldm(shR + 9 * a), ld3(shSize + 3 * a), shRadius[a], ld3(shX + 3 * b),
                      ldm(shR + 9 * b), ld3(shSize + 3 * b), shRadius[b], sink);
  } else {
    int pl = a, bx = b;
    xf T;
    T.R = ldm(shR + 9 * bx);
    T.t = ld3(shX + 3 * bx);
    n = collideBoxPlane(T, ld3(shSize + 3 * bx), shRadius[bx], ld3(shSize + 3 * pl), shRadius[pl], sink);
  }
  pairCount[i] = n;
}

// Sphere-tree pairs: one warp per pair, explicit stack in shared memory, lanes test the children of the
// node being descended in parallel (fan-out <= 30 in the shipped .sph files; larger fan-outs loop).
// The visited set is exactly the reference recursion's (collideSphereTrees :975-1009 "descend the larger
// sphere", collideSphereTreeAndPlane :793-826, collideBoxAndSphereTree :837-850); only the visiting order
// differs, which does not affect the contact set.
#define TREE_STACK 1024
#define WARPS_PER_BLOCK 4

struct TreeCtx {
  const int* shType;
  const int* shRoot;
  const double* shSize;
  const double* shRadius;
  const double* shP;
  const double* shX;
  const double* shR;
  const double* ndC;
  const double* ndR;
  const int* ndFirst;
  const int* ndCount;
};

#define SPLIT_TARGET 512  // node pairs waiting at which a tree x tree pair is handed over to one warp per node pair
struct TreeTasks {
  unsigned long long* val;  // node pair (node of tree 1 << 32 | node of tree 2) a task starts from
  int* pair;                // candidate pair it belongs to
  int* count;               // contacts it found (count pass)
  int* prefix;              // exclusive scan of count, [nTasks + 1]
  int* counter;             // tasks allocated so far
  int cap;
  int target;               // node pairs waiting at which a pair is handed over (SPLIT_TARGET; a huge value = never)
  int* pairStart;           // per candidate pair: its tasks [pairStart, pairStart + pairN) and the contacts of its own prefix walk
  int* pairN;
  int* pairPrefix;
};

// One step of the tree x tree descent (collideSphereTrees :975-1009): the children of the larger sphere of (n1, n2) -
// of node2 when node1 is a leaf, of node1 when node2 is - are tested against the other node by the lanes; overlapping
// pairs that are not leaf x leaf are pushed at stack[sp..], leaf x leaf contacts are counted (EMIT: written at base + count).
template <bool EMIT>
__device__ __forceinline__ void treeTreeStep(unsigned long long top, const xf& Ta, const xf& Tb, const TreeCtx& C, const HitOut& H, long long base,
                                             int& count, unsigned long long* stack, int& sp, int* __restrict__ overflowFlag,
                                             const int stackCap = TREE_STACK) {
  const int lane = threadIdx.x & 31;
  int n1 = (int)(top >> 32), n2 = (int)(top & 0xffffffffu);
  bool l1 = C.ndFirst[n1] < 0, l2 = C.ndFirst[n2] < 0;
  double r1p = C.ndR[n1], r2p = C.ndR[n2];
  bool descend2;  // iterate the children of node2?
  if (l1) descend2 = true; else if (l2) descend2 = false; else descend2 = (r1p <= r2p);
  int dn = descend2 ? n2 : n1;
  int first = C.ndFirst[dn], cnt = C.ndCount[dn];
  for (int k0 = 0; k0 < cnt; k0 += 32) {
    int k = k0 + lane;
    bool push = false, emit = false;
    int m1 = n1, m2 = n2;
    d3 c1, c2;
    double r1 = 0, r2 = 0, dist = 0;
    if (k < cnt) {
      if (descend2) m2 = first + k; else m1 = first + k;
      c1 = xfP(Ta, ld3(C.ndC + 3 * m1));
      c2 = xfP(Tb, ld3(C.ndC + 3 * m2));
      r1 = C.ndR[m1];
      r2 = C.ndR[m2];
      if (vdist2(c1, c2) < (r1 + r2) * (r1 + r2)) {
        if (C.ndFirst[m1] < 0 && C.ndFirst[m2] < 0) {
          dist = vdist(c1, c2);
          emit = dist < r2 + r1;
        } else {
          push = true;
        }
      }
    }
    unsigned pm = __ballot_sync(0xffffffffu, push), em = __ballot_sync(0xffffffffu, emit);
    if (push) {
      int pos = sp + __popc(pm & ((1u << lane) - 1));
      if (pos < stackCap) stack[pos] = ((unsigned long long)(unsigned)m1 << 32) | (unsigned)m2;
    }
    sp += __popc(pm);
    if (sp > stackCap) { if (lane == 0) *overflowFlag = 1; sp = stackCap; }
    if (emit && EMIT) {
      int pos = count + __popc(em & ((1u << lane) - 1));
      double dbc = r2 + r1;
      double alpha = (r1 - r2 + dist) / (2 * dist);
      d3 p((1 - alpha) * c1.x + alpha * c2.x, (1 - alpha) * c1.y + alpha * c2.y, (1 - alpha) * c1.z + alpha * c2.z);
      d3 nrm = vnormalize(vsub(c2, c1));
      writeHit(H, base + pos, p, nrm, dist - dbc, 0, m1, m2, -1);
    }
    count += __popc(em);
    __syncwarp();
  }
}

template <bool EMIT>
__device__ __forceinline__ void narrowTreePair(int i, unsigned long long* stack, const unsigned long long* __restrict__ pairVal,
                                               const int* __restrict__ pairType, const int* __restrict__ pairSlot,
                                               const TreeCtx& C, const HitOut& H, int* __restrict__ pairCountOrCap,
                                               int* __restrict__ overflowFlag, const TreeTasks& T) {
  int lane = threadIdx.x & 31;
  int pt = pairType[i];
  int a = (int)(pairVal[i] >> 32), b = (int)(pairVal[i] & 0xffffffffu);
  long long base = EMIT ? pairSlot[i] : 0;
  int count = 0;  // warp-uniform
  int sp = 0;     // warp-uniform stack pointer

  if (pt == PT_TREETREE) {
    xf Ta, Tb;
    Ta.R = ldm(C.shR + 9 * a); Ta.t = ld3(C.shX + 3 * a);
    Tb.R = ldm(C.shR + 9 * b); Tb.t = ld3(C.shX + 3 * b);
    int ra = C.shRoot[a], rb = C.shRoot[b];
    // the root pair is tested like any other pair
    {
      d3 c1 = xfP(Ta, ld3(C.ndC + 3 * ra)), c2 = xfP(Tb, ld3(C.ndC + 3 * rb));
      double r1 = C.ndR[ra], r2 = C.ndR[rb];
      bool hit = vdist2(c1, c2) < (r1 + r2) * (r1 + r2);
      if (hit) {
        if (C.ndFirst[ra] < 0 && C.ndFirst[rb] < 0) {
          double dist = vdist(c1, c2), dbc = r2 + r1;
          if (dist < dbc) {
            if (EMIT && lane == 0) {
              double alpha = (r1 - r2 + dist) / (2 * dist);
              d3 p((1 - alpha) * c1.x + alpha * c2.x, (1 - alpha) * c1.y + alpha * c2.y, (1 - alpha) * c1.z + alpha * c2.z);
              d3 nrm = vnormalize(vsub(c2, c1));
              writeHit(H, base + count, p, nrm, dist - dbc, 0, ra, rb, -1);
            }
            count++;
          }
        } else {
          if (lane == 0) stack[0] = ((unsigned long long)(unsigned)ra << 32) | (unsigned)rb;
          sp = 1;
        }
      }
    }
    __syncwarp();
    // Breadth first while the frontier is small: a pair of meshes pressed together holds 10^3 - 10^4 leaf x leaf contacts,
    // and one warp walking all of it was the tail of the whole narrowphase.  Once SPLIT_TARGET node pairs are waiting, they
    // are handed to the task list (one warp each, k_tree_tasks) and this warp is done; small pairs finish right here.  The
    // contacts of a pair are emitted in the order [this prefix][task 0][task 1]... - the same in the count and in the emit
    // pass, which walk identically.
    int head = 0;
    while (head < sp && sp - head < T.target && sp <= TREE_STACK - 64) {
      unsigned long long top = stack[head];
      head++;
      __syncwarp();
      treeTreeStep<EMIT>(top, Ta, Tb, C, H, base, count, stack, sp, overflowFlag);
    }
    int nT = sp - head;
    if (!EMIT) {
      int start = 0;
      if (nT > 0) {
        if (lane == 0) start = atomicAdd(T.counter, nT);
        start = __shfl_sync(0xffffffffu, start, 0);
        if (start + nT <= T.cap)
          for (int j = lane; j < nT; j += 32) { T.val[start + j] = stack[head + j]; T.pair[start + j] = i; }
      }
      if (lane == 0) { T.pairStart[i] = start; T.pairN[i] = nT; }
    }
  } else {
    // single tree against a plane or a box: stack of node indices
    int tr = (pt == PT_TREEPLANE) ? a : b;
    int ot = (pt == PT_TREEPLANE) ? b : a;
    xf Tt, To;
    Tt.R = ldm(C.shR + 9 * tr); Tt.t = ld3(C.shX + 3 * tr);
    To.R = ldm(C.shR + 9 * ot); To.t = ld3(C.shX + 3 * ot);
    d3 pn, pp, bsize;
    double pd = 0, brad = 0;
    bool plane = pt == PT_TREEPLANE;
    if (plane) { pn = ld3(C.shSize + 3 * ot); pd = C.shRadius[ot]; pp = ld3(C.shP + 3 * ot); }
    else { bsize = ld3(C.shSize + 3 * ot); brad = C.shRadius[ot]; }
    // virtual parent whose single child is the root
    int first = C.shRoot[tr], cnt = 1;
    bool firstRound = true;
    while (firstRound || sp > 0) {
      if (!firstRound) {
        int nd = (int)stack[sp - 1];
        sp--;
        __syncwarp();
        first = C.ndFirst[nd];
        cnt = C.ndCount[nd];
      }
      firstRound = false;
      for (int k0 = 0; k0 < cnt; k0 += 32) {
        int k = k0 + lane;
        bool push = false, emit = false;
        int m = first + k;
        d3 c;
        double r = 0, dpl = 0;
        Hit h;
        if (k < cnt) {
          c = xfP(Tt, ld3(C.ndC + 3 * m));
          r = C.ndR[m];
          bool leafNode = C.ndFirst[m] < 0;
          if (plane) {
            dpl = pn.x * c.x + pn.y * c.y + pn.z * c.z + pd - r;
            if (dpl < 0) { if (leafNode) emit = true; else push = true; }
          } else {
            if (overlapBoxSphere(To, bsize, brad, c, r)) {
              if (leafNode) emit = collideBoxSphere(To, bsize, c, r, &h) > 0; else push = true;
            }
          }
        }
        unsigned pm = __ballot_sync(0xffffffffu, push), em = __ballot_sync(0xffffffffu, emit);
        if (push) {
          int pos = sp + __popc(pm & ((1u << lane) - 1));
          if (pos < TREE_STACK) stack[pos] = (unsigned long long)(unsigned)m;
        }
        sp += __popc(pm);
        if (sp > TREE_STACK) { if (lane == 0) *overflowFlag = 1; sp = TREE_STACK; }
        if (emit && EMIT) {
          int pos = count + __popc(em & ((1u << lane) - 1));
          if (plane) {
            // p = c - n (n . (c - p0)), contact normal = -n   (CollisionProcessor.java:809-813)
            d3 nv = vsub(c, pp);
            double val = vdot(nv, pn);
            nv = vscale(val, pn);
            d3 cw = vsub(c, nv);
            writeHit(H, base + pos, cw, vscale(-1, pn), dpl, 0, m, AM3D_BV_PLANE_DUMMY, -1);
          } else {
            writeHit(H, base + pos, h.pos, h.normal, h.violation, 0, AM3D_BV_NULL, AM3D_BV_NULL, m);
          }
        }
        count += __popc(em);
        __syncwarp();
      }
    }
  }
  if (!EMIT) {
    if (lane == 0) {
      T.pairPrefix[i] = count;        // (k_tree_paircap adds the tasks' contacts to pairCap)
      pairCountOrCap[i] = count;
      if (pt != PT_TREETREE) T.pairN[i] = 0;
    }
  } else if (lane == 0) {
    int s0 = T.pairStart[i], n = T.pairN[i];
    pairCountOrCap[i] = count + (n > 0 ? T.prefix[s0 + n] - T.prefix[s0] : 0);
  }
}

// One warp per entry of the tree-pair list written by k_pair_classify (scenes of boxes with a few meshes have a
// million pairs and a handful of tree pairs); the grid is fixed and the warps stride over the list, whose length
// lives on the device.
template <bool EMIT>
__global__ void k_narrow_tree(const int* __restrict__ treeList, const int* __restrict__ treeCount,
                              const unsigned long long* __restrict__ pairVal, const int* __restrict__ pairType,
                              const int* __restrict__ pairSlot, TreeCtx C, HitOut H, int* __restrict__ pairCountOrCap,
                              int* __restrict__ overflowFlag, TreeTasks T) {
  __shared__ unsigned long long stackMem[WARPS_PER_BLOCK][TREE_STACK];
  int warp = threadIdx.x >> 5;
  int n = *treeCount;
  for (int t = blockIdx.x * WARPS_PER_BLOCK + warp; t < n; t += gridDim.x * WARPS_PER_BLOCK) {
    narrowTreePair<EMIT>(treeList[t], stackMem[warp], pairVal, pairType, pairSlot, C, H, pairCountOrCap, overflowFlag, T);
    __syncwarp();
  }
}

// One warp per task: the depth-first descent from one node pair of a tree x tree candidate pair.  A depth-first stack
// holds at most (levels of both trees) x (fan-out) entries - a small one keeps the shared memory per warp at 4 KB and the
// occupancy at what the registers allow (the walk is a chain of dependent loads: it lives on resident warps).
#define TASK_STACK 512
template <bool EMIT>
__global__ void k_tree_tasks(int nTasks, const unsigned long long* __restrict__ pairVal, const int* __restrict__ pairSlot, TreeCtx C, HitOut H,
                             int* __restrict__ overflowFlag, TreeTasks T) {
  __shared__ unsigned long long stackMem[WARPS_PER_BLOCK][TASK_STACK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* stack = stackMem[warp];
  for (int t = blockIdx.x * WARPS_PER_BLOCK + warp; t < nTasks; t += gridDim.x * WARPS_PER_BLOCK) {
    const int i = T.pair[t];
    const int a = (int)(pairVal[i] >> 32), b = (int)(pairVal[i] & 0xffffffffu);
    xf Ta, Tb;
    Ta.R = ldm(C.shR + 9 * a); Ta.t = ld3(C.shX + 3 * a);
    Tb.R = ldm(C.shR + 9 * b); Tb.t = ld3(C.shX + 3 * b);
    long long base = 0;
    if (EMIT) base = (long long)pairSlot[i] + T.pairPrefix[i] + (T.prefix[t] - T.prefix[T.pairStart[i]]);
    int count = 0, sp = 1;
    if (lane == 0) stack[0] = T.val[t];
    __syncwarp();
    while (sp > 0) {
      unsigned long long top = stack[sp - 1];
      sp--;
      __syncwarp();
      treeTreeStep<EMIT>(top, Ta, Tb, C, H, base, count, stack, sp, overflowFlag, TASK_STACK);
    }
    if (!EMIT && lane == 0) T.count[t] = count;
    __syncwarp();
  }
}
// capacity of a tree pair = the contacts of its prefix walk + those of its tasks
__global__ void k_tree_paircap(const int* __restrict__ treeList, const int* __restrict__ treeCount, TreeTasks T, int* __restrict__ pairCap) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *treeCount) return;
  int i = treeList[t];
  int n = T.pairN[i];
  if (n > 0) pairCap[i] = T.pairPrefix[i] + (T.prefix[T.pairStart[i] + n] - T.prefix[T.pairStart[i]]);
}

// ------------------------------------------------------------------------------------------------
// Contact.set (Contact.java:159-233) fused with the slot -> canonical-order compaction
// ------------------------------------------------------------------------------------------------
struct ContactOut {
  int *b1, *b2, *s1, *s2, *bv1, *bv2, *info, *leaf, *state, *isNew;
  unsigned long long *key0, *key1;
  double *pW, *nW, *t1W, *t2W, *pB1, *nB1, *t1B1, *t2B1, *viol, *prevViol, *lam, *lamWarm;
};

// owner[c] = the candidate pair that emitted contact c (canonical order): lets k_contact_set run one thread per
// CONTACT, so that every store of the ~50 it does per contact is coalesced across the warp
__global__ void k_contact_owner(int np, const int* __restrict__ pairCount, const int* __restrict__ pairOut, int* __restrict__ owner) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  int n = pairCount[i], out = pairOut[i];
  for (int k = 0; k < n; k++) owner[out + k] = i;
}

__global__ void k_contact_set(int nc, const int* __restrict__ owner, const int* __restrict__ bShapeFirst,
                              const unsigned long long* __restrict__ pairVal, const int* __restrict__ pairSlot,
                              const int* __restrict__ pairOut,
                              const int* __restrict__ shBody, const double* __restrict__ x, const double* __restrict__ R,
                              const double* __restrict__ hitPos, const double* __restrict__ hitNrm,
                              const double* __restrict__ hitViol, const int* __restrict__ hitMeta, ContactOut O) {
  int out = blockIdx.x * blockDim.x + threadIdx.x;
  if (out >= nc) return;
  int i = owner[out];
  int s1 = (int)(pairVal[i] >> 32), s2 = (int)(pairVal[i] & 0xffffffffu);
  int b1 = shBody[s1], b2 = shBody[s2];  // composite parts report their parent (Contact.java:169-180)
  xf T1;
  T1.R = ldm(R + 9 * b1);
  T1.t = ld3(x + 3 * b1);
  unsigned long long k0 = pairKey0(shBody, bShapeFirst, s1, s2);
  long long slot = (long long)pairSlot[i] + (out - pairOut[i]);
  {
    int4 m = reinterpret_cast<const int4*>(hitMeta)[slot];
    d3 p = ld3(hitPos + 3 * slot), nW = ld3(hitNrm + 3 * slot);
    double anx = fabs(nW.x), any = fabs(nW.y), anz = fabs(nW.z);
    d3 t1;
    if (anx < any && anx < anz) t1 = d3(1, 0, 0);
    else if (any < anz) t1 = d3(0, 1, 0);
    else t1 = d3(0, 0, 1);
    d3 t2 = vcross(nW, t1);
    t2 = vnormalize(t2);
    t1 = vcross(t2, nW);
    O.b1[out] = b1; O.b2[out] = b2; O.s1[out] = s1; O.s2[out] = s2;
    O.bv1[out] = m.y; O.bv2[out] = m.z; O.info[out] = m.x; O.leaf[out] = m.w;
    O.state[out] = AM3D_CS_CLEAR;
    O.isNew[out] = 1;
    int bvLo = b1 < b2 ? m.y : m.z, bvHi = b1 < b2 ? m.z : m.y;
    O.key0[out] = k0;
    O.key1[out] = ((unsigned long long)(bvLo + 2) << 36) | ((unsigned long long)(bvHi + 2) << 8) | (unsigned long long)m.x;
    st3(O.pW + 3 * out, p); st3(O.nW + 3 * out, nW); st3(O.t1W + 3 * out, t1); st3(O.t2W + 3 * out, t2);
    st3(O.pB1 + 3 * out, xfInvP(T1, p));
    st3(O.nB1 + 3 * out, mtransformT(T1.R, nW));
    st3(O.t1B1 + 3 * out, mtransformT(T1.R, t1));
    st3(O.t2B1 + 3 * out, mtransformT(T1.R, t2));
    O.viol[out] = hitViol[slot];
    O.prevViol[out] = 0.0;
    O.lam[3 * out] = 0; O.lam[3 * out + 1] = 0; O.lam[3 * out + 2] = 0;
    O.lamWarm[3 * out] = 0; O.lamWarm[3 * out + 1] = 0; O.lamWarm[3 * out + 2] = 0;
  }
}
