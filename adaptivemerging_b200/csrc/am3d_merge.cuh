// Kernel group 5 of the north star: the adaptive-merging machinery.
//
//   merge      Merging.merge (Merging.java:73-163) + RigidCollection.addBody/addCollection (:128-178,
//              477-514): the sequential greedy union over a HashSet of body pairs becomes a lock-free
//              union-find over top-level entities; collection mass properties (COM, mass-weighted v and
//              omega, parallel-axis inertia, bounding box) are recomputed from the members by one block per
//              changed collection with a fixed-shape tree reduction (deterministic).  The partition is the
//              reference's (connected components of the mergeable pairs plus every pair whose two bodies
//              end up in one collection, RigidCollection.addIncompleteContacts :987-998); floating-point
//              summation order differs from the reference's incremental updates (SURVEY.md "Hard parts" 7).
//   unmerge    Merging.unmerge/unmergeSelectedBpcs (:215-374): cut flags per internal body pair
//              (BodyPairContact.checkContactsState :255-266, checkMotionMetricForUnmerging :200-205),
//              connected components of what is left (union-find over the members), pieces smaller than
//              n/2+1 leave, the rest keeps the collection (RigidCollection.removeBodies :686-753).
//   metric     BodyPairContact.accumulateForUnmerging :127-145.
#pragma once
#include "am3d_ctx.h"
#include "am3d_math.cuh"
#include "am3d_step.cuh"

// ------------------------------------------------------------------------------------------------
// union-find (roots are the smallest index of their set)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ufFind(int* uf, int i) {
  volatile int* u = uf;
  while (true) {
    int p = u[i];
    if (p == i) return i;
    int g = u[p];
    if (g != p) u[i] = g;  // path halving (benign race: only ever moves a node closer to its root)
    i = p;
  }
}
__device__ __forceinline__ void ufUnite(int* uf, int a, int b) {
  while (true) {
    a = ufFind(uf, a);
    b = ufFind(uf, b);
    if (a == b) return;
    if (a > b) { int t = a; a = b; b = t; }
    int old = atomicCAS(uf + b, b, a);
    if (old == b) return;
  }
}
__global__ void k_uf_init(int n, int* __restrict__ uf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) uf[i] = i;
}
__global__ void k_uf_flatten(int n, int* __restrict__ uf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) uf[i] = ufFind(uf, i);
}

// ------------------------------------------------------------------------------------------------
// merge
// ------------------------------------------------------------------------------------------------
struct MergeParamsD {
  int mergePinned, stableContact, breathe, accum;
  double thrMerge, thrBreath;
};

// BodyPairContact.checkMergeCondition :151-177 on every live external body pair
__global__ void k_merge_flag(int nbp, const int* __restrict__ alive, const int* __restrict__ bb1,
                             const int* __restrict__ bb2, const int* __restrict__ bstart, const int* __restrict__ bcount,
                             const double* __restrict__ lam, const double* __restrict__ viol,
                             const double* __restrict__ prevViol, const int* __restrict__ flags,
                             const int* __restrict__ parent, const double* __restrict__ mh, const int* __restrict__ sh,
                             const int* __restrict__ nm, const int* __restrict__ nst, MergeParamsD P,
                             int* __restrict__ flag, int* __restrict__ uf, int* __restrict__ nFlagged) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbp) return;
  flag[b] = 0;
  if (!alive[b]) return;
  int l1 = bb1[b], l2 = bb2[b];
  bool ok = true;
  if ((flags[l1] & AM3D_F_SLEEPING) && (flags[l2] & AM3D_F_SLEEPING)) {
    ok = true;
  } else {
    if (P.breathe) {
      for (int i = bstart[b]; i < bstart[b] + bcount[b]; i++) {
        if (fabs(lam[3 * i]) > 1e-14 && fabs(prevViol[i] - viol[i]) > P.thrBreath) { ok = false; break; }
      }
    }
    if (ok && !P.mergePinned && ((flags[l1] & AM3D_F_PINNED) || (flags[l2] & AM3D_F_PINNED))) ok = false;
    if (ok && parent[l1] >= 0 && parent[l1] == parent[l2]) ok = false;
    if (ok) {
      if (nm[b] == P.accum) {
        for (int k = 0; k < nm[b]; k++) if (mh[4 * b + k] > P.thrMerge) { ok = false; break; }
      } else ok = false;
    }
    if (ok && P.stableContact) {
      if (nst[b] == P.accum) {
        for (int k = 0; k < nst[b]; k++) if (sh[4 * b + k] == AM3D_CS_ONEDGE) { ok = false; break; }
      } else ok = false;
    }
  }
  if (!ok) return;
  flag[b] = 1;
  atomicAdd(nFlagged, 1);
  int a = parent[l1] >= 0 ? parent[l1] : l1, c = parent[l2] >= 0 ? parent[l2] : l2;
  ufUnite(uf, a, c);
}

// per top-level entity: size of its component (in entities) and the best pre-existing collection of it
__global__ void k_merge_census(int ns, int nb, const int* __restrict__ collAlive, const int* __restrict__ parent,
                               const int* __restrict__ uf, const int* __restrict__ collCount,
                               const long long* __restrict__ stamp, int* __restrict__ compEnt,
                               unsigned long long* __restrict__ compBest) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ns) return;
  if (e < nb ? parent[e] >= 0 : !collAlive[e - nb]) return;
  int r = uf[e];
  atomicAdd(compEnt + r, 1);
  if (e >= nb) {
    // most members wins, ties to the older collection (smaller stamp); low bits carry the slot
    unsigned long long key = ((unsigned long long)collCount[e - nb] << 40) | (unsigned long long)(0xffffffffffULL - (unsigned long long)stamp[e]);
    atomicMax(compBest + r, key);
  }
}
// ---- Merging.merge as the SEQUENCE the reference runs (Merging.java:73-163) ----------------------------------------
// Which collection object survives a merge, and where a new one enters RigidBodySystem.bodies, depends on the order
// in which the mergeable pairs are visited: two free bodies found a collection (appended to the list), a free body
// joins the collection of its partner, of two collections the one with more bodies absorbs the other (ties: body2's,
// Merging.java:113-129).  The reference visits a HashSet; the oracle canonicalises that to ascending (lo, hi) pair
// order, which is the order of the body-pair table.  Components of the merge graph are independent, so one thread
// replays the pairs of one component in that order on a private union-find (dpar/dsize/dident, indexed by top-level
// entity); components with more than `maxEdges` mergeable pairs fall back to "largest collection survives, ties to
// the oldest" (k_merge_census) - at those sizes neither the reference nor the oracle can serve as a comparison.
#define MS_FREE 0x7fffffff
#define MS_FALLBACK 0x7ffffffe
__global__ void k_mseq_init(int ns, int nb, const int* __restrict__ collAlive, const int* __restrict__ collCount,
                            int* __restrict__ dpar, int* __restrict__ dsize, int* __restrict__ dident, int* __restrict__ survivor) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ns) return;
  dpar[e] = e;
  bool coll = e >= nb && collAlive[e - nb];
  dsize[e] = coll ? collCount[e - nb] : 1;
  dident[e] = coll ? e - nb : MS_FREE;
  survivor[e] = MS_FALLBACK;
}
// the mergeable pairs, in table order (= ascending (lo, hi) unless an unmerge appended pairs this step: then the host
// sorts them by pair key first)
__global__ void k_mseq_list(int nbp, const int* __restrict__ mflag, const int* __restrict__ scan,
                            const unsigned long long* __restrict__ bkey, int* __restrict__ list, unsigned long long* __restrict__ lkey) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbp || !mflag[b]) return;
  int o = scan[b];
  list[o] = b;
  lkey[o] = bkey[b];
}
// per mergeable pair (position k in visiting order): component root as sort key, k as value
__global__ void k_mseq_keys(int n, const int* __restrict__ list, const int* __restrict__ bb1, const int* __restrict__ parent,
                            const int* __restrict__ uf, unsigned int* __restrict__ key, int* __restrict__ val) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int l1 = bb1[list[k]];
  int e = parent[l1] >= 0 ? parent[l1] : l1;
  key[k] = (unsigned)uf[e];
  val[k] = k;
}
__global__ void k_seg_heads(int n, const unsigned int* __restrict__ key, int* __restrict__ head) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}
__global__ void k_seg_fill(int n, const int* __restrict__ head, const int* __restrict__ scan, int* __restrict__ segStart) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (head[i]) segStart[scan[i]] = i;
  if (i == n - 1) segStart[scan[i] + head[i]] = n;
}
__device__ __forceinline__ int msFind(int* dpar, int e) {
  while (true) {
    int p = dpar[e];
    if (p == e) return e;
    int g = dpar[p];
    dpar[e] = g;
    e = g;
  }
}
__global__ void k_mseq_run(int nseg, const int* __restrict__ segStart, const unsigned int* __restrict__ keySorted,
                           const int* __restrict__ valSorted, const int* __restrict__ list, const int* __restrict__ bb1, const int* __restrict__ bb2,
                           const int* __restrict__ parent, int maxEdges, int* __restrict__ dpar, int* __restrict__ dsize,
                           int* __restrict__ dident, int* __restrict__ survivor) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  int k0 = segStart[s], k1 = segStart[s + 1];
  int root = (int)keySorted[k0];
  if (k1 - k0 > maxEdges) return;  // survivor stays MS_FALLBACK
  for (int k = k0; k < k1; k++) {
    int seq = valSorted[k];  // position in the visiting order
    int b = list[seq];
    int l1 = bb1[b], l2 = bb2[b];
    int e1 = parent[l1] >= 0 ? parent[l1] : l1, e2 = parent[l2] >= 0 ? parent[l2] : l2;
    int r1 = msFind(dpar, e1), r2 = msFind(dpar, e2);
    if (r1 == r2) continue;  // already together (RigidCollection.addIncompleteContacts made the pair internal)
    int i1 = dident[r1], i2 = dident[r2];
    int keep, other, ident, size;
    if (i1 == MS_FREE && i2 == MS_FREE) { keep = r1; other = r2; ident = -1 - seq; size = 2; }             // :100-110
    else if (i1 != MS_FREE && i2 != MS_FREE) {                                                                  // :111-129
      if (dsize[r1] > dsize[r2]) { keep = r1; other = r2; } else { keep = r2; other = r1; }
      ident = dident[keep]; size = dsize[r1] + dsize[r2];
    } else if (i1 != MS_FREE) { keep = r1; other = r2; ident = i1; size = dsize[r1] + 1; }                    // :130-141
    else { keep = r2; other = r1; ident = i2; size = dsize[r2] + 1; }                                           // :142-152
    dpar[other] = keep;
    dident[keep] = ident;
    dsize[keep] = size;
  }
  survivor[root] = dident[msFind(dpar, root)];
}
// components of >1 entity whose surviving collection is a new one need a fresh slot
__global__ void k_merge_neednew(int ns, const int* __restrict__ uf, const int* __restrict__ compEnt,
                                const unsigned long long* __restrict__ compBest, const int* __restrict__ survivor,
                                int* __restrict__ needNew) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ns) return;
  bool comp = uf[r] == r && compEnt[r] > 1;
  int sv = survivor[r];
  needNew[r] = (comp && (sv == MS_FALLBACK ? compBest[r] == 0ULL : sv < 0)) ? 1 : 0;
}
__global__ void k_free_slots(int nc, const int* __restrict__ collAlive, int* __restrict__ isFree) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nc) isFree[c] = collAlive[c] ? 0 : 1;
}
__global__ void k_free_list(int nc, const int* __restrict__ isFree, const int* __restrict__ scan, int* __restrict__ list) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nc && isFree[c]) list[scan[c]] = c;
}
// target collection slot of every merging component root
__global__ void k_merge_target(int ns, int nb, const int* __restrict__ uf, const int* __restrict__ compEnt,
                               const unsigned long long* __restrict__ compBest, const int* __restrict__ needNew,
                               const int* __restrict__ newScan, const int* __restrict__ freeList,
                               const long long* __restrict__ stamp, const int* __restrict__ collAlive,
                               const int* __restrict__ survivor, int* __restrict__ target) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ns) return;
  target[r] = -1;
  if (uf[r] != r || compEnt[r] <= 1) return;
  if (needNew[r]) target[r] = freeList[newScan[r]];
  else if (survivor[r] != MS_FALLBACK) target[r] = survivor[r];
  else {
    // recover the slot from the winning (count, stamp) key: find the collection of this component with that stamp
    target[r] = -2;  // resolved by k_merge_target2
  }
}
__global__ void k_merge_target2(int nc, int nb, const int* __restrict__ collAlive, const int* __restrict__ uf,
                                const int* __restrict__ collCount, const long long* __restrict__ stamp,
                                const unsigned long long* __restrict__ compBest, int* __restrict__ target) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc || !collAlive[c]) return;
  int r = uf[nb + c];
  unsigned long long key = ((unsigned long long)collCount[c] << 40) | (unsigned long long)(0xffffffffffULL - (unsigned long long)stamp[nb + c]);
  if (target[r] == -2 && compBest[r] == key) target[r] = c;
}
// re-parent leaves, retire absorbed collections, start new ones
__global__ void k_merge_apply(int ns, int nb, int* __restrict__ collAlive, int* __restrict__ parent,
                              const int* __restrict__ uf, const int* __restrict__ compEnt, const int* __restrict__ target,
                              const int* __restrict__ needNew, const int* __restrict__ newScan, int* __restrict__ flags,
                              long long* __restrict__ stamp, long long stampBase, int* __restrict__ collMode,
                              int* __restrict__ collFlagAcc, int* __restrict__ metricCount) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ns) return;
  if (e < nb) {
    int t = parent[e] >= 0 ? parent[e] : e;
    int r = uf[t];
    if (compEnt[r] <= 1) return;
    int tg = target[r];
    if (parent[e] < 0) {
      // a free body joins: RigidCollection.updateCollectionState :459-469
      int f = flags[e];
      if (f & AM3D_F_SLEEPING) { atomicOr(collFlagAcc + tg, AM3D_F_SLEEPING); atomicAnd(flags + e, ~AM3D_F_SLEEPING); }
    }
    parent[e] = nb + tg;
  } else {
    int c = e - nb;
    if (!collAlive[c]) return;
    int r = uf[e];
    if (compEnt[r] <= 1) return;
    int tg = target[r];
    if (flags[e] & AM3D_F_SLEEPING) atomicOr(collFlagAcc + tg, AM3D_F_SLEEPING);
    if (tg != c) collAlive[c] = 0;  // absorbed (Merging.java:117,127)
  }
}
__global__ void k_merge_newcolls(int ns, int nb, const int* __restrict__ uf, const int* __restrict__ compEnt,
                                 const int* __restrict__ target, const int* __restrict__ needNew,
                                 const int* __restrict__ newScan, const int* __restrict__ survivor, int nFlagged,
                                 int* __restrict__ collAlive, int* __restrict__ flags,
                                 long long* __restrict__ stamp, long long stampBase, int* __restrict__ collMode,
                                 int* __restrict__ metricCount) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ns) return;
  if (uf[r] != r || compEnt[r] <= 1) return;
  int tg = target[r];
  collMode[tg] = needNew[r] ? 3 : 1;
  if (needNew[r]) {
    collAlive[tg] = 1;
    flags[nb + tg] = 0;
    // appended to RigidBodySystem.bodies when the pair that founded it was visited
    int sv = survivor[r];
    stamp[nb + tg] = sv == MS_FALLBACK ? stampBase + nFlagged + newScan[r] : stampBase + (-1 - sv);
    metricCount[nb + tg] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// membership lists (CSR): leaves sorted by (collection slot, leaf id)
// ------------------------------------------------------------------------------------------------
__global__ void k_member_keys(int nb, int nc, const int* __restrict__ parent, unsigned int* __restrict__ key,
                              int* __restrict__ val, int* __restrict__ collCount) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  int p = parent[i];
  key[i] = p >= 0 ? (unsigned)(p - nb) : (unsigned)nc;
  val[i] = i;
  if (p >= 0) atomicAdd(collCount + (p - nb), 1);
}

// ------------------------------------------------------------------------------------------------
// collection mass properties: one block per changed collection
// ------------------------------------------------------------------------------------------------
#define CR_THREADS 256
__device__ __forceinline__ void blockSum(double* sh, double& v) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = CR_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
    __syncthreads();
  }
  v = sh[0];
  __syncthreads();
}
__device__ __forceinline__ void blockMinMax(double* sh, double& v, bool isMax) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = CR_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = isMax ? fmax(sh[threadIdx.x], sh[threadIdx.x + s]) : fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  v = sh[0];
  __syncthreads();
}

// mode 1: new or grown collection (RigidCollection.addBodyInternalMethod :477-514, applied to all members at once)
// mode 2: shrunk collection (RigidCollection.removeBodies :686-753): v and omega are kept
__global__ void __launch_bounds__(CR_THREADS)
k_coll_recompute(int nc, int nb, const int* __restrict__ changedList, const int* __restrict__ collMode,
                 const int* __restrict__ collStart, const int* __restrict__ collCount, const int* __restrict__ members,
                 const int* __restrict__ collFlagAcc, double* __restrict__ x, double* __restrict__ R, double* __restrict__ v,
                 double* __restrict__ w, double* __restrict__ mass, double* __restrict__ minv, double* __restrict__ mA,
                 double* __restrict__ mA0, double* __restrict__ jinv, double* __restrict__ jinv0, int* __restrict__ flags,
                 double* __restrict__ bbB, int* __restrict__ bbCount, double* __restrict__ B2CR, double* __restrict__ B2Ct) {
  __shared__ double sh[CR_THREADS];
  __shared__ int shFlag;
  int c = changedList[blockIdx.x];
  int mode = collMode[c];
  int cs = nb + c;
  int s0 = collStart[c], n = collCount[c];
  if (threadIdx.x == 0) shFlag = 0;
  __syncthreads();
  double M = 0, mx[3] = {0, 0, 0}, mv[3] = {0, 0, 0}, mw[3] = {0, 0, 0};
  double lo[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
  double hi[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
  int pinned = 0, anyBox = 0;
  for (int k = threadIdx.x; k < n; k += CR_THREADS) {
    int b = members[s0 + k];
    double m = mass[b];
    int f = flags[b];
    if (f & AM3D_F_PINNED) pinned = 1;
    d3 xb = ld3(x + 3 * b), vb = ld3(v + 3 * b), wb = ld3(w + 3 * b);
    M += m;
    mx[0] += m * xb.x; mx[1] += m * xb.y; mx[2] += m * xb.z;
    mv[0] += m * vb.x; mv[1] += m * vb.y; mv[2] += m * vb.z;
    mw[0] += m * wb.x; mw[1] += m * wb.y; mw[2] += m * wb.z;
    int nbb = bbCount[b];
    if (nbb) anyBox = 1;
    xf T;
    T.R = ldm(R + 9 * b);
    T.t = xb;
    for (int q = 0; q < nbb; q++) {
      d3 p = xfP(T, ld3(bbB + 24 * b + 3 * q));
      lo[0] = fmin(lo[0], p.x); lo[1] = fmin(lo[1], p.y); lo[2] = fmin(lo[2], p.z);
      hi[0] = fmax(hi[0], p.x); hi[1] = fmax(hi[1], p.y); hi[2] = fmax(hi[2], p.z);
    }
  }
  if (pinned) atomicOr(&shFlag, 1);
  if (anyBox) atomicOr(&shFlag, 2);
  blockSum(sh, M);
  for (int k = 0; k < 3; k++) { blockSum(sh, mx[k]); blockSum(sh, mv[k]); blockSum(sh, mw[k]); blockMinMax(sh, lo[k], false); blockMinMax(sh, hi[k], true); }
  __syncthreads();
  pinned = shFlag & 1;
  anyBox = (shFlag & 2) ? 1 : 0;
  d3 com;
  if (pinned) com = ld3(x + 3 * cs);  // a pinned collection keeps whatever origin it has; it never moves
  else com = d3(mx[0] / M, mx[1] / M, mx[2] / M);
  if (pinned && mode == 3) {
    // a brand-new pinned collection takes the origin of its first non-plane member
    // (RigidCollection(body1, body2) :55-78: set(body) copies x before the pinned branch freezes it)
    int b0 = members[s0];
    for (int k = 0; k < n; k++) if (bbCount[members[s0 + k]]) { b0 = members[s0 + k]; break; }
    com = ld3(x + 3 * b0);
  }
  // inertia about the COM: sum of member inertias moved by the parallel-axis term (RigidCollection.getOp :809-830)
  double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (!pinned) {
    for (int k = threadIdx.x; k < n; k += CR_THREADS) {
      int b = members[s0 + k];
      double m = mass[b];
      double dx = x[3 * b] - com.x, dy = x[3 * b + 1] - com.y, dz = x[3 * b + 2] - com.z;
      double x2 = dx * dx, y2 = dy * dy, z2 = dz * dz;
      double op[9] = {y2 + z2, -dx * dy, -dx * dz, -dy * dx, x2 + z2, -dy * dz, -dz * dx, -dz * dy, x2 + y2};
      for (int q = 0; q < 9; q++) J[q] += mA[9 * b + q] + m * op[q];
    }
    for (int q = 0; q < 9; q++) blockSum(sh, J[q]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int f = flags[cs] & ~(AM3D_F_PINNED);
    if (mode == 1 || mode == 3) f |= collFlagAcc[c];
    if (pinned) f |= AM3D_F_PINNED;
    flags[cs] = f;
    st3(x + 3 * cs, com);
    stm(R + 9 * cs, midentity());  // theta.setIdentity() (RigidCollection.updateTheta :544-582, removeBodies :705)
    if (pinned) {
      st3(v + 3 * cs, d3()); st3(w + 3 * cs, d3());
      mass[cs] = 0; minv[cs] = 0;
      stm(mA + 9 * cs, mzero()); stm(mA0 + 9 * cs, mzero()); stm(jinv + 9 * cs, mzero()); stm(jinv0 + 9 * cs, mzero());
    } else {
      if (mode == 1 || mode == 3) {
        st3(v + 3 * cs, d3(mv[0] / M, mv[1] / M, mv[2] / M));
        st3(w + 3 * cs, d3(mw[0] / M, mw[1] / M, mw[2] / M));
      }
      mass[cs] = M;
      minv[cs] = 1. / M;
      m3 Jm, Ji;
      for (int q = 0; q < 9; q++) Jm.m[q] = J[q];
      if (!minvert(Jm, Ji)) Ji = mzero();
      stm(mA + 9 * cs, Jm); stm(mA0 + 9 * cs, Jm); stm(jinv + 9 * cs, Ji); stm(jinv0 + 9 * cs, Ji);
    }
    if (anyBox) {
      d3 mn(lo[0] - com.x, lo[1] - com.y, lo[2] - com.z), mxx(hi[0] - com.x, hi[1] - com.y, hi[2] - com.z);
      double* B = bbB + 24 * cs;
      st3(B + 12, mn);
      st3(B + 15, d3(mn.x, mxx.y, mxx.z));
      st3(B + 18, d3(mxx.x, mn.y, mxx.z));
      st3(B + 21, d3(mxx.x, mxx.y, mn.z));
      st3(B + 0, mxx);
      st3(B + 3, d3(mxx.x, mn.y, mn.z));
      st3(B + 6, d3(mn.x, mxx.y, mn.z));
      st3(B + 9, d3(mn.x, mn.y, mxx.z));
      bbCount[cs] = 8;
    } else {
      bbCount[cs] = 0;
    }
  }
  // RigidCollection.updateBodiesTransformations :587-590 with theta = I: B2C = (R_i, x_i - com)
  for (int k = threadIdx.x; k < n; k += CR_THREADS) {
    int b = members[s0 + k];
    for (int q = 0; q < 9; q++) B2CR[9 * b + q] = R[9 * b + q];
    B2Ct[3 * b] = x[3 * b] - com.x; B2Ct[3 * b + 1] = x[3 * b + 1] - com.y; B2Ct[3 * b + 2] = x[3 * b + 2] - com.z;
  }
}
__global__ void k_changed_flags(int nc, const int* __restrict__ collAlive, const int* __restrict__ collMode, int* __restrict__ flag) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nc) flag[c] = (collAlive[c] && collMode[c] != 0) ? 1 : 0;
}
__global__ void k_changed_list(int nc, const int* __restrict__ flag, const int* __restrict__ scan, int* __restrict__ list) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nc && flag[c]) list[scan[c]] = c;
}

// ------------------------------------------------------------------------------------------------
// internalising body pairs (RigidCollection.addToInternalContact :836-847, addIncompleteContacts :987-998)
// ------------------------------------------------------------------------------------------------
__global__ void k_int_flag(int nbp, int* __restrict__ alive, const int* __restrict__ bb1, const int* __restrict__ bb2,
                           const int* __restrict__ parent, const int* __restrict__ nActive, int* __restrict__ toInt,
                           int* __restrict__ toIntContacts) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbp) return;
  int t = 0;
  if (alive[b]) {
    int p1 = parent[bb1[b]], p2 = parent[bb2[b]];
    if (p1 >= 0 && p1 == p2) t = 1;
  }
  toInt[b] = t;
  toIntContacts[b] = t ? nActive[b] : 0;
  if (t) alive[b] = 0;  // leaves the external set (Merging.java:159-162)
}

struct ContactPtrs {
  int *b1, *b2, *s1, *s2, *bv1, *bv2, *info, *leaf, *bpc, *state, *isNew;
  unsigned long long *key0, *key1;
  double *pW, *nW, *t1W, *t2W, *pB1, *nB1, *t1B1, *t2B1, *viol, *prevViol, *lam, *lamWarm;
};
__device__ __forceinline__ void copyContact(const ContactPtrs& S, int i, const ContactPtrs& D, int o) {
  D.b1[o] = S.b1[i]; D.b2[o] = S.b2[i]; D.s1[o] = S.s1[i]; D.s2[o] = S.s2[i]; D.bv1[o] = S.bv1[i]; D.bv2[o] = S.bv2[i];
  D.info[o] = S.info[i]; D.leaf[o] = S.leaf[i]; D.state[o] = S.state[i]; D.key0[o] = S.key0[i]; D.key1[o] = S.key1[i];
  for (int k = 0; k < 3; k++) {
    D.pW[3 * o + k] = S.pW[3 * i + k]; D.nW[3 * o + k] = S.nW[3 * i + k]; D.t1W[3 * o + k] = S.t1W[3 * i + k]; D.t2W[3 * o + k] = S.t2W[3 * i + k];
    D.pB1[3 * o + k] = S.pB1[3 * i + k]; D.nB1[3 * o + k] = S.nB1[3 * i + k]; D.t1B1[3 * o + k] = S.t1B1[3 * i + k]; D.t2B1[3 * o + k] = S.t2B1[3 * i + k];
    D.lam[3 * o + k] = S.lam[3 * i + k];
  }
  D.viol[o] = S.viol[i]; D.prevViol[o] = S.prevViol[i];
}
// one thread per body pair that becomes internal: append it and its ACTIVE contacts to the internal tables
__global__ void k_int_copy(int nbp, const int* __restrict__ toInt, const int* __restrict__ bscan, const int* __restrict__ cscan,
                           const unsigned long long* __restrict__ bkey, const int* __restrict__ bb1, const int* __restrict__ bb2,
                           const int* __restrict__ bstart, const int* __restrict__ bcount, ContactPtrs S, int ibBase, int icBase,
                           unsigned long long* __restrict__ ikey, int* __restrict__ ib1, int* __restrict__ ib2,
                           int* __restrict__ istart, int* __restrict__ icount, int* __restrict__ ialive,
                           int* __restrict__ inMetric, int* __restrict__ icut, ContactPtrs D) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbp || !toInt[b]) return;
  int ob = ibBase + bscan[b];
  int oc = icBase + cscan[b];
  ikey[ob] = bkey[b]; ib1[ob] = bb1[b]; ib2[ob] = bb2[b]; istart[ob] = oc; ialive[ob] = 1; inMetric[ob] = 0; icut[ob] = 0;
  int n = 0;
  for (int i = bstart[b]; i < bstart[b] + bcount[b]; i++) {
    if (fabs(S.lam[3 * i]) > 1e-14) {
      copyContact(S, i, D, oc + n);
      D.bpc[oc + n] = ob;
      D.isNew[oc + n] = 0;
      for (int k = 0; k < 3; k++) D.lamWarm[3 * (oc + n) + k] = 0;
      n++;
    }
  }
  icount[ob] = n;
}

// ------------------------------------------------------------------------------------------------
// unmerge
// ------------------------------------------------------------------------------------------------
// accumulateForUnmerging (BodyPairContact.java:127-145) + cut decision (Merging.java:231-244)
__global__ void k_unm_flag(int nib, const int* __restrict__ ialive, const int* __restrict__ ib1, const int* __restrict__ ib2,
                           const int* __restrict__ istart, const int* __restrict__ icount, const int* __restrict__ cstate,
                           const int* __restrict__ parent, const int* __restrict__ flags, const double* __restrict__ x,
                           const double* __restrict__ R, const double* __restrict__ v, const double* __restrict__ w,
                           const double* __restrict__ bbB, const int* __restrict__ bbCount, int nb, double thrUnmerge,
                           int accumUnmerge, int unmNormal, int unmFriction, double posDt, int* __restrict__ inMetric, int* __restrict__ icut,
                           int* __restrict__ collCuts, int* __restrict__ total) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nib) return;
  icut[b] = 0;
  if (!ialive[b]) return;
  int l1 = ib1[b], l2 = ib2[b];
  int p = parent[l1];
  if (p < 0 || p != parent[l2]) return;
  if (flags[p] & AM3D_F_SLEEPING) return;
  double metric = pairMetric(l1, l2, flags, x, R, v, w, bbB, bbCount, posDt);
  int nm = inMetric[b];
  if (metric > thrUnmerge) { if (nm < 1000000) nm++; } else nm = 0;
  inMetric[b] = nm;
  bool cut = false;
  for (int i = istart[b]; i < istart[b] + icount[b]; i++) {
    int st = cstate[i];
    if (st == AM3D_CS_BROKEN && unmNormal) { cut = true; break; }
    if (st == AM3D_CS_ONEDGE && unmFriction) { cut = true; break; }
  }
  if (!cut && nm >= accumUnmerge && icount[b] < 3) cut = true;
  if (cut) {
    icut[b] = 1;
    atomicAdd(collCuts + (p - nb), 1);
    atomicAdd(total, 1);
  }
}
// union the members over the uncut internal pairs of every collection that has a cut
__global__ void k_unm_union(int nib, const int* __restrict__ ialive, const int* __restrict__ icut, const int* __restrict__ ib1,
                            const int* __restrict__ ib2, const int* __restrict__ parent, const int* __restrict__ collCuts, int nb,
                            int* __restrict__ uf) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nib || !ialive[b] || icut[b]) return;
  int l1 = ib1[b], l2 = ib2[b];
  int p = parent[l1];
  if (p < 0 || p != parent[l2] || collCuts[p - nb] == 0) return;
  ufUnite(uf, l1, l2);
}
__global__ void k_unm_census(int nb, const int* __restrict__ parent, const int* __restrict__ collCuts, const int* __restrict__ uf,
                             int* __restrict__ compSize, int* __restrict__ collNComp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  int p = parent[i];
  if (p < 0 || collCuts[p - nb] == 0) return;
  atomicAdd(compSize + uf[i], 1);
  if (uf[i] == i) atomicAdd(collNComp + (p - nb), 1);
}
// roots of the pieces that leave: singles become free bodies, larger pieces need a new collection slot
__global__ void k_unm_roots(int nb, const int* __restrict__ parent, const int* __restrict__ collCuts, const int* __restrict__ uf,
                            const int* __restrict__ compSize, const int* __restrict__ collNComp, const int* __restrict__ collCount,
                            int* __restrict__ leaves /* per leaf root: 0 stays, 1 leaves */, int* __restrict__ needNew,
                            int* __restrict__ freed, int* __restrict__ collKeeps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  leaves[i] = 0; needNew[i] = 0; freed[i] = 0;
  int p = parent[i];
  if (p < 0 || collCuts[p - nb] == 0 || uf[i] != i) return;
  int c = p - nb;
  if (collNComp[c] <= 1) return;
  int n = collCount[c];
  if (compSize[i] < n / 2 + 1) {
    leaves[i] = 1;
    if (compSize[i] > 1) needNew[i] = 1; else freed[i] = 1;
  } else {
    collKeeps[c] = 1;
  }
}
// rank of the leaving pieces in the order the reference appends them to RigidBodySystem.bodies: collections in list
// order (stamp), pieces of one collection by their smallest member
__global__ void k_unm_rank_keys(int nb, const int* __restrict__ leaves, const int* __restrict__ parent,
                                const long long* __restrict__ stamp, unsigned long long* __restrict__ key, int* __restrict__ val) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  val[i] = i;
  key[i] = leaves[i] ? (((unsigned long long)stamp[parent[i]] << 24) | (unsigned long long)i) : 0xffffffffffffffffULL;
}
__global__ void k_unm_rank_scatter(int n, const int* __restrict__ sortedVal, int* __restrict__ rank) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) rank[sortedVal[p]] = p;
}
__global__ void k_unm_apply(int nb, int* __restrict__ parent, const int* __restrict__ collCuts, const int* __restrict__ uf,
                            const int* __restrict__ leaves, const int* __restrict__ needNew, const int* __restrict__ newScan,
                            const int* __restrict__ freeList, const int* __restrict__ orderScan /* scan over leaving roots */,
                            const double* __restrict__ x, double* __restrict__ v, double* __restrict__ w, double* __restrict__ dv,
                            long long* __restrict__ stamp, long long stampBase, int* __restrict__ collAlive,
                            int* __restrict__ collMode, int* __restrict__ flags, int* __restrict__ metricCount) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  int p = parent[i];
  if (p < 0 || collCuts[p - nb] == 0) return;
  int r = uf[i];
  if (!leaves[r]) {
    collMode[p - nb] = 2;  // shrunk (if anything left at all; harmless when the collection is unchanged)
    return;
  }
  // RigidCollection.unmergeBody :944-953
  d3 rr = vsub(ld3(x + 3 * i), ld3(x + 3 * p));
  d3 om = ld3(w + 3 * p);
  st3(v + 3 * i, vadd(ld3(v + 3 * p), vcross(om, rr)));
  st3(w + 3 * i, om);
  for (int k = 0; k < DVS; k++) dv[DVS * (size_t)i + k] = 0;
  if (needNew[r]) {
    int slot = freeList[newScan[r]];
    parent[i] = nb + slot;
    if (i == r) {
      collAlive[slot] = 1;
      collMode[slot] = 3;
      flags[nb + slot] = 0;
      stamp[nb + slot] = stampBase + orderScan[r];
      metricCount[nb + slot] = 0;
    }
  } else {
    parent[i] = -1;
    stamp[i] = stampBase + orderScan[r];  // re-enters RigidBodySystem.bodies at the end (Merging.java:269)
  }
}
__global__ void k_unm_retire(int nc, const int* __restrict__ collCuts, const int* __restrict__ collNComp,
                             const int* __restrict__ collKeeps, int* __restrict__ collAlive, int* __restrict__ collMode) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  if (collCuts[c] == 0 || !collAlive[c]) return;
  if (collNComp[c] <= 1) { collMode[c] = 0; return; }  // not actually split: nothing changes
  if (!collKeeps[c] && collMode[c] != 3) { collAlive[c] = 0; collMode[c] = 0; }
}
// cut pairs whose bodies are still together stay internal, the others go back to the external set
__global__ void k_unm_reext_flag(int nib, const int* __restrict__ ialive, int* __restrict__ icut, const int* __restrict__ ib1,
                                 const int* __restrict__ ib2, const int* __restrict__ icount, const int* __restrict__ parent,
                                 int* __restrict__ toExt, int* __restrict__ toExtContacts) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nib) return;
  int t = 0;
  if (ialive[b] && icut[b]) {
    int p1 = parent[ib1[b]], p2 = parent[ib2[b]];
    if (p1 >= 0 && p1 == p2) icut[b] = 0;  // reconnect (Merging.java:344-346)
    else t = 1;
  }
  toExt[b] = t;
  toExtContacts[b] = t ? icount[b] : 0;
}
__global__ void k_unm_reext_copy(int nib, const int* __restrict__ toExt, const int* __restrict__ bscan, const int* __restrict__ cscan,
                                 const unsigned long long* __restrict__ ikey, const int* __restrict__ ib1, const int* __restrict__ ib2,
                                 const int* __restrict__ istart, const int* __restrict__ icount, int* __restrict__ ialive,
                                 ContactPtrs S, int bpBase, int cBase, unsigned long long* __restrict__ bkey, int* __restrict__ bb1,
                                 int* __restrict__ bb2, int* __restrict__ bstart, int* __restrict__ bcount, int* __restrict__ balive,
                                 int* __restrict__ bnActive, int* __restrict__ bnm, int* __restrict__ bnst, ContactPtrs D,
                                 const double* __restrict__ x, const double* __restrict__ R) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nib || !toExt[b]) return;
  int ob = bpBase + bscan[b], oc = cBase + cscan[b];
  bkey[ob] = ikey[b]; bb1[ob] = ib1[b]; bb2[ob] = ib2[b]; bstart[ob] = oc; bcount[ob] = icount[b]; balive[ob] = 1;
  bnActive[ob] = 0; bnm[ob] = 0; bnst[ob] = 0;  // histories cleared (Merging.java:357-358)
  ialive[b] = 0;
  for (int k = 0; k < icount[b]; k++) {
    int i = istart[b] + k, o = oc + k;
    copyContact(S, i, D, o);
    D.bpc[o] = ob;
    D.isNew[o] = 0;
    for (int q = 0; q < 3; q++) D.lamWarm[3 * o + q] = S.lam[3 * i + q];  // Merging.java:351-353
    // world frame for the export / non-collection Jacobian path
    int l1 = S.b1[i];
    xf T;
    T.R = ldm(R + 9 * l1);
    T.t = ld3(x + 3 * l1);
    st3(D.pW + 3 * o, xfP(T, ld3(S.pB1 + 3 * i)));
    st3(D.nW + 3 * o, mtransform(T.R, ld3(S.nB1 + 3 * i)));
    st3(D.t1W + 3 * o, mtransform(T.R, ld3(S.t1B1 + 3 * i)));
    st3(D.t2W + 3 * o, mtransform(T.R, ld3(S.t2B1 + 3 * i)));
  }
}

// stable compaction of the internal tables after pairs left
__global__ void k_ibp_compact_flag(int nib, const int* __restrict__ ialive, const int* __restrict__ icount, int* __restrict__ keep,
                                   int* __restrict__ keepContacts) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nib) return;
  keep[b] = ialive[b] ? 1 : 0;
  keepContacts[b] = ialive[b] ? icount[b] : 0;
}
__global__ void k_ibp_compact(int nib, const int* __restrict__ keep, const int* __restrict__ bscan, const int* __restrict__ cscan,
                              const unsigned long long* __restrict__ ikey, const int* __restrict__ ib1, const int* __restrict__ ib2,
                              const int* __restrict__ istart, const int* __restrict__ icount, const int* __restrict__ inMetric,
                              ContactPtrs S, unsigned long long* __restrict__ okey, int* __restrict__ ob1, int* __restrict__ ob2,
                              int* __restrict__ ostart, int* __restrict__ ocount, int* __restrict__ oalive, int* __restrict__ onMetric,
                              int* __restrict__ ocut, ContactPtrs D) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nib || !keep[b]) return;
  int ob = bscan[b], oc = cscan[b];
  okey[ob] = ikey[b]; ob1[ob] = ib1[b]; ob2[ob] = ib2[b]; ostart[ob] = oc; ocount[ob] = icount[b]; oalive[ob] = 1;
  onMetric[ob] = inMetric[b]; ocut[ob] = 0;
  for (int k = 0; k < icount[b]; k++) {
    copyContact(S, istart[b] + k, D, oc + k);
    D.bpc[oc + k] = ob;
    D.isNew[oc + k] = 0;
    for (int q = 0; q < 3; q++) D.lamWarm[3 * (oc + k) + q] = S.lamWarm[3 * (istart[b] + k) + q];
  }
}

// ------------------------------------------------------------------------------------------------
// single sweep support (CollisionProcessor.updateInCollections :232-303)
// ------------------------------------------------------------------------------------------------
// group list of the sweep: external pairs [0,nExt) followed by internal pairs [nExt, nExt+nInt); internal
// pairs of sleeping collections take no part (count 0)
__global__ void k_sweep_groups(int nExt, int nInt, const int* __restrict__ ecb1, const int* __restrict__ ecb2,
                               const int* __restrict__ ecount, const int* __restrict__ estart, const int* __restrict__ ib1,
                               const int* __restrict__ icb1, const int* __restrict__ icb2, const int* __restrict__ icount,
                               const int* __restrict__ istart, const int* __restrict__ ialive, const int* __restrict__ parent,
                               const int* __restrict__ flags, int* __restrict__ gb1, int* __restrict__ gb2,
                               int* __restrict__ gcount, int* __restrict__ gstart, int* __restrict__ gasleep) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nExt + nInt) return;
  // The bodies of a group are taken in the orientation of ITS CONTACTS (Contact.body1/body2 of this step), which can
  // differ from the orientation the BodyPairContact object was created with once list positions changed.
  if (g < nExt) {
    int s = estart[g];
    gb1[g] = ecb1[s]; gb2[g] = ecb2[s]; gcount[g] = ecount[g]; gstart[g] = s; gasleep[g] = 0;
  } else {
    // internal pair of a collection.  Pairs of a SLEEPING collection take part in the sweep only if the breadth-first
    // walk from the new contacts reaches them (CollisionProcessor.java:389-394 walks body.bodyPairContacts without
    // looking at sleeping flags; only the "missing bpc from collections" pass :405-415 skips sleeping collections)
    int b = g - nExt;
    int s = istart[b];
    bool live = ialive[b] && icount[b] > 0;
    int p = live ? parent[ib1[b]] : -1;
    live = live && p >= 0;
    gb1[g] = live ? icb1[s] : 0; gb2[g] = live ? icb2[s] : 0; gstart[g] = s;
    gcount[g] = live ? icount[b] : 0;
    gasleep[g] = (live && (flags[p] & AM3D_F_SLEEPING)) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// getOrganizedContacts (CollisionProcessor.java:346-441): breadth-first layers over the body-pair graph, starting from
// the pairs that hold a contact that is new this time step; two pairs are neighbours when they share a leaf body
// (pinned ones included: body.bodyPairContacts of the plane lists every pair resting on it).
// ------------------------------------------------------------------------------------------------
#define BFS_INF 0x7f7f7f7f
__global__ void k_bfs_seed(int nc, const int* __restrict__ isNew, const int* __restrict__ cbpc, int* __restrict__ grpLayer) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  if (isNew[i] && cbpc[i] >= 0) grpLayer[cbpc[i]] = 0;
}
// body pairs of the bodies the user interacted with (RigidBody.picked, set by the mouse tools) come right after the
// pairs with new contacts (CollisionProcessor.java:361-383): class 1; the breadth-first layers follow as classes 2, 3, ...
__global__ void k_bfs_seed_picked(int ng, const int* __restrict__ gb1, const int* __restrict__ gb2, const int* __restrict__ gcount,
                                  const int* __restrict__ picked, int* __restrict__ grpLayer) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng || gcount[g] == 0 || grpLayer[g] == 0) return;
  if (picked[gb1[g]] || picked[gb2[g]]) grpLayer[g] = 1;
}
// cooperative: one grid barrier per layer.  round[3] = "something was reached" flags, rotating; round[3] = deepest layer
__global__ void __launch_bounds__(256)
k_bfs_layers(int ng, const int* __restrict__ gb1, const int* __restrict__ gb2, const int* __restrict__ gcount,
             int* __restrict__ grpLayer, int* __restrict__ bodyLevel, int* __restrict__ round) {
  cg::grid_group grid = cg::this_grid();
  int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  for (int g = tid; g < ng; g += stride)
    if (gcount[g] > 0 && grpLayer[g] <= 1) { atomicMin(bodyLevel + gb1[g], 1); atomicMin(bodyLevel + gb2[g], 1); }  // seeds: classes 0 and 1
  grid.sync();
  for (int L = 2;; L++) {
    if (tid == 0) round[(L + 1) % 3] = 0;
    bool any = false;
    for (int g = tid; g < ng; g += stride) {
      if (gcount[g] > 0 && grpLayer[g] == BFS_INF) {
        int a = gb1[g], b = gb2[g];
        if (__ldcg(bodyLevel + a) < L || __ldcg(bodyLevel + b) < L) {
          grpLayer[g] = L;
          atomicMin(bodyLevel + a, L);
          atomicMin(bodyLevel + b, L);
          any = true;
        }
      }
    }
    if (any) round[L % 3] = 1;
    grid.sync();
    if (__ldcg(round + L % 3) == 0) {
      if (tid == 0) round[3] = L - 1;
      break;
    }
  }
}
// The same walk for a context of many independent scenes: the layers of a scene depend on that scene alone, so a CTA takes a
// scene and separates its layers with __syncthreads() instead of a grid barrier - a pile is 50 - 150 layers deep, and a
// grid barrier plus a pass over the groups of ALL scenes per layer was 1.6 ms per sweep at 512 scenes (0.02 ms when no
// contact is new).  sceneStart / sceneList: the groups bucketed by scene (k_scene_count / k_scene_fill; their order inside a
// bucket is irrelevant: within a layer nobody reads what the layer writes).  round[3] = deepest layer over all scenes.
__global__ void k_scene_count(int ng, const int* __restrict__ gb1, const int* __restrict__ bodyScene, int* __restrict__ cnt) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < ng) atomicAdd(cnt + bodyScene[gb1[g]], 1);
}
__global__ void k_scene_fill(int ng, const int* __restrict__ gb1, const int* __restrict__ bodyScene, const int* __restrict__ start,
                             int* __restrict__ cursor, int* __restrict__ list) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  int s = bodyScene[gb1[g]];
  list[start[s] + atomicAdd(cursor + s, 1)] = g;
}
__global__ void __launch_bounds__(256)
k_bfs_scenes(const int* __restrict__ sceneStart, const int* __restrict__ sceneList, const int* __restrict__ gb1, const int* __restrict__ gb2,
             const int* __restrict__ gcount, int* __restrict__ grpLayer, int* __restrict__ bodyLevel, int* __restrict__ round) {
  const int g0 = sceneStart[blockIdx.x], g1 = sceneStart[blockIdx.x + 1];
  for (int i = g0 + threadIdx.x; i < g1; i += blockDim.x) {
    int g = sceneList[i];
    if (gcount[g] > 0 && grpLayer[g] <= 1) { atomicMin(bodyLevel + gb1[g], 1); atomicMin(bodyLevel + gb2[g], 1); }  // seeds: classes 0 and 1
  }
  __syncthreads();
  for (int L = 2;; L++) {
    int any = 0;
    for (int i = g0 + threadIdx.x; i < g1; i += blockDim.x) {
      int g = sceneList[i];
      if (gcount[g] > 0 && grpLayer[g] == BFS_INF) {
        int a = gb1[g], b = gb2[g];
        if (__ldcg(bodyLevel + a) < L || __ldcg(bodyLevel + b) < L) {
          grpLayer[g] = L;
          atomicMin(bodyLevel + a, L);
          atomicMin(bodyLevel + b, L);
          any = 1;
        }
      }
    }
    if (!__syncthreads_or(any)) {
      if (threadIdx.x == 0) atomicMax(round + 3, L - 1);
      break;
    }
  }
}
// pairs the walk did not reach: external ones follow the last layer, internal ones of awake collections come last,
// internal ones of sleeping collections stay out of the sweep
__global__ void k_bfs_finalize(int ng, int nExt, const int* __restrict__ gasleep, const int* __restrict__ round,
                               int* __restrict__ grpLayer, int* __restrict__ gcount) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  if (grpLayer[g] != BFS_INF) return;
  int deepest = round[3];
  if (g < nExt) grpLayer[g] = deepest + 1;
  else {
    grpLayer[g] = deepest + 2;
    if (gasleep[g]) gcount[g] = 0;
  }
}
// A body pair with hundreds of contacts (sphere trees) is cut into CHUNKS of at most `ch` consecutive contacts, and the
// chunks - not the pair - are what gets coloured and scheduled: the chunks of one pair share both bodies, so they land
// in different phases, but the pairs of OTHER bodies interleave with them.  The sweep then costs about the largest
// per-body contact load instead of (number of giant-holding colours) x (longest pair).  Any interleaving is a
// Gauss-Seidel sequence; the oracle replays the one chosen here.
__global__ void k_chunk_count(int np, const int* __restrict__ pcount, int ch, int* __restrict__ n) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < np) n[g] = pcount[g] > ch ? (pcount[g] + ch - 1) / ch : 1;
}
__global__ void k_chunk_expand(int np, int ch, const int* __restrict__ first, const int* __restrict__ pb1, const int* __restrict__ pb2,
                               const int* __restrict__ pcount, const int* __restrict__ pstart, const int* __restrict__ player,
                               int* __restrict__ gb1, int* __restrict__ gb2, int* __restrict__ gcount, int* __restrict__ gstart,
                               int* __restrict__ glayer, int* __restrict__ glead) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= np) return;
  int e0 = first[g], e1 = first[g + 1];
  int cnt = pcount[g], st = pstart[g];
  for (int e = e0; e < e1; e++) {
    int off = (e - e0) * ch;
    gb1[e] = pb1[g]; gb2[e] = pb2[g];
    gstart[e] = st + off;
    gcount[e] = (e1 - e0 == 1) ? cnt : min(ch, cnt - off);
    if (player) glayer[e] = player[g];
    glead[e] = e == e0;
  }
}
// organize_contacts = false: external pairs first, then the internal pairs of awake collections
__global__ void k_plain_layers(int ng, int nExt, const int* __restrict__ gasleep, int* __restrict__ grpLayer, int* __restrict__ gcount) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  grpLayer[g] = g < nExt ? 0 : 1;
  if (g >= nExt && gasleep[g]) gcount[g] = 0;
}
// after the sweep: sub-bodies of awake collections advance their velocities with their own deltaV
// (CollisionProcessor.java:289-297), then every deltaV is zeroed for the full solve (:299)
__global__ void k_sweep_finish(int ns, int nb, const int* __restrict__ parent, const int* __restrict__ flags,
                               const double* __restrict__ minv, const double* __restrict__ jinv, const double* __restrict__ force,
                               const double* __restrict__ torque, double* __restrict__ dv, double* __restrict__ v,
                               double* __restrict__ w, double dt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  if (i < nb) {
    int p = parent[i];
    if (p >= 0 && !(flags[p] & AM3D_F_SLEEPING) && !(flags[i] & AM3D_F_PINNED)) {
      d3 vv = vscaleAdd(dt * minv[i], ld3(force + 3 * i), ld3(v + 3 * i));
      vv = vadd(vv, ld3(dv + DVS * (size_t)i));
      d3 dom = mtransform(ldm(jinv + 9 * i), ld3(torque + 3 * i));
      dom = vscale(dt, dom);
      d3 ww = vadd(ld3(w + 3 * i), dom);
      ww = vadd(ww, ld3(dv + DVS * (size_t)i + 3));
      st3(v + 3 * i, vv);
      st3(w + 3 * i, ww);
    }
  }
  for (int k = 0; k < DVS; k++) dv[DVS * (size_t)i + k] = 0;
}
// the re-clear + re-apply of external forces on top-level bodies once a merge event happened (:138-142)
__global__ void k_reclear_top(int ns, int nb, const int* __restrict__ alive, const int* __restrict__ parent,
                              const double* __restrict__ mass, double* __restrict__ force, double* __restrict__ torque,
                              double* __restrict__ dv, int useGravity, double gx, double gy) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  if (i >= nb ? !alive[i - nb] : parent[i] >= 0) return;
  d3 f(0, 0, 0);
  if (useGravity) f = vadd(f, vscale(-mass[i], d3(gx, gy, 0)));
  st3(force + 3 * i, f);
  st3(torque + 3 * i, d3(0, 0, 0));
  for (int k = 0; k < DVS; k++) dv[DVS * (size_t)i + k] = 0;
}
// collections carry their members along (RigidCollection.updateBodiesPositionAndTransformations :898-909,
// applyVelocitiesToBodies :914-918)
// RigidCollection.clearBodies :100-105 (postStabilization): EVERY collection, pinned and sleeping ones too, hands its
// velocity to its members
__global__ void k_members_take_velocity(int nb, const int* __restrict__ parent, const double* __restrict__ x, double* __restrict__ v,
                                        double* __restrict__ w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  int p = parent[i];
  if (p < 0) return;
  d3 r = vsub(ld3(x + 3 * i), ld3(x + 3 * p));
  d3 om = ld3(w + 3 * p);
  st3(v + 3 * i, vadd(ld3(v + 3 * p), vcross(om, r)));
  st3(w + 3 * i, om);
}
__global__ void k_members_follow(int nb, const int* __restrict__ parent, const int* __restrict__ flags, int pushVel, int pushPos,
                                 double* __restrict__ x, double* __restrict__ R, double* __restrict__ v, double* __restrict__ w,
                                 const double* __restrict__ B2CR, const double* __restrict__ B2Ct, const double* __restrict__ jinv0,
                                 const double* __restrict__ mA0, double* __restrict__ jinv, double* __restrict__ mA) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  int p = parent[i];
  if (p < 0) return;
  int fp = flags[p];
  if (fp & (AM3D_F_PINNED | AM3D_F_SLEEPING)) return;  // only advanced collections move their members
  if (pushPos) {
    xf C, B;
    C.R = ldm(R + 9 * p); C.t = ld3(x + 3 * p);
    B.R = ldm(B2CR + 9 * i); B.t = ld3(B2Ct + 3 * i);
    xf T = xfMul(C, B);
    stm(R + 9 * i, T.R);
    st3(x + 3 * i, T.t);
    stm(jinv + 9 * i, rm0rt(T.R, ldm(jinv0 + 9 * i)));
    stm(mA + 9 * i, rm0rt(T.R, ldm(mA0 + 9 * i)));
  }
  if (pushVel) {
    d3 r = vsub(ld3(x + 3 * i), ld3(x + 3 * p));
    d3 om = ld3(w + 3 * p);
    st3(v + 3 * i, vadd(ld3(v + 3 * p), vcross(om, r)));
    st3(w + 3 * i, om);
  }
}
