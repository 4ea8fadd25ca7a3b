// Kernels 3, 4 and 6 of the north star plus the per-step bookkeeping kernels:
//   external forces (RigidBodySystem.java:234-290, Spring.java:153-209), body-pair bookkeeping
//   (CollisionProcessor.java:145-226), warm start (:481-665), sleeping (Sleeping.java:49-139),
//   contact frame / Jacobian / b / D assembly (Contact.java:235-354), graph-coloured projected
//   Gauss-Seidel (PGS.java:73-256), integration (RigidBody.java:382-441), motion metric
//   (MotionMetricProcessor.java:39-73, BodyPairContact.java:83-121).
#pragma once
#include "am3d_ctx.h"
#include <cooperative_groups.h>

#include "am3d_math.cuh"
namespace cg = cooperative_groups;

#define BLK 256
static inline int nblk(long long n, int b = BLK) { return (int)((n + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// forces
// ------------------------------------------------------------------------------------------------
// clearBodies :190-202 + RigidCollection.clearBodies :100-105 + applyGravityForce :276-290
__global__ void k_clear_gravity(int ns, int nb, const int* __restrict__ alive, const int* __restrict__ parent,
                                const double* __restrict__ mass, const double* __restrict__ x, double* __restrict__ v,
                                double* __restrict__ w, double* __restrict__ force, double* __restrict__ torque,
                                double* __restrict__ dv, int useGravity, double gx, double gy) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  if (i >= nb && !alive[i - nb]) return;
  if (i < nb && parent[i] >= 0) {
    // RigidCollection.applyVelocitiesTo :925-937
    int c = parent[i];
    d3 r = vsub(ld3(x + 3 * i), ld3(x + 3 * c));
    d3 om = ld3(w + 3 * c);
    d3 wxr = vcross(om, r);
    st3(v + 3 * i, vadd(ld3(v + 3 * c), wxr));
    st3(w + 3 * i, om);
  }
  d3 f(0, 0, 0);
  if (useGravity) {
    double m = mass[i];
    d3 tmp = vscale(-m, d3(gx, gy, 0));
    f = vadd(f, tmp);
  }
  st3(force + 3 * i, f);
  st3(torque + 3 * i, d3(0, 0, 0));
#pragma unroll
  for (int k = 0; k < DVS; k++) dv[DVS * (size_t)i + k] = 0.0;
}

// useCoriolis (RigidBodySystem.java:212-229 gyroscopicStabilization, :295-304 applyCoriolis, RigidBody.java:348-354):
// awake unpinned top-level bodies get massAngular -= dt^2 * Lhat * jinv * Lhat (L = massAngular * omega), then every
// unpinned body (members of collections included) gets torque -= (massAngular * omega) x omega
__global__ void k_coriolis(int ns, int nb, const int* __restrict__ alive, const int* __restrict__ parent, const int* __restrict__ flags,
                           const double* __restrict__ w, const double* __restrict__ jinv, double* __restrict__ mA,
                           double* __restrict__ torque, double dt, int topOnly) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  if (i >= nb && !alive[i - nb]) return;
  bool top = i >= nb || parent[i] < 0;
  if (topOnly && !top) return;
  int f = flags[i];
  d3 om = ld3(w + 3 * i);
  m3 M = ldm(mA + 9 * i);
  if (top && !(f & (AM3D_F_PINNED | AM3D_F_SLEEPING))) {
    d3 L = mtransform(M, om);
    m3 Lh;
    Lh.m[0] = 0.; Lh.m[1] = -L.z; Lh.m[2] = L.y;
    Lh.m[3] = L.z; Lh.m[4] = 0.; Lh.m[5] = -L.x;
    Lh.m[6] = -L.y; Lh.m[7] = L.x; Lh.m[8] = 0.;
#pragma unroll
    for (int k = 0; k < 9; k++) Lh.m[k] = Lh.m[k] * dt;
    m3 T = mmul(Lh, ldm(jinv + 9 * i));
    T = mmul(T, Lh);
#pragma unroll
    for (int k = 0; k < 9; k++) { T.m[k] = T.m[k] * -1.; M.m[k] = M.m[k] + T.m[k]; }
    stm(mA + 9 * i, M);
  }
  if (!(f & AM3D_F_PINNED)) {
    d3 tmp = mtransform(M, om);
    d3 tmp2 = vcross(tmp, om);
    st3(torque + 3 * i, vsub(ld3(torque + 3 * i), tmp2));
  }
}

__device__ __forceinline__ d3 spatialVelocity(const double* x, const double* v, const double* w, int b, const d3& pW) {
  d3 tmp = vsub(pW, ld3(x + 3 * b));
  d3 r = vcross(ld3(w + 3 * b), tmp);
  return vadd(r, ld3(v + 3 * b));
}
__device__ __forceinline__ void applyForceW(double* force, double* torque, const double* x, int b, const d3& pW, const d3& fW,
                                            bool atomic) {
  d3 tmp = vsub(pW, ld3(x + 3 * b));
  d3 tq = vcross(tmp, fW);
  if (!atomic) {
    st3(force + 3 * b, vadd(ld3(force + 3 * b), fW));
    st3(torque + 3 * b, vadd(ld3(torque + 3 * b), tq));
  } else {
    atomicAdd(force + 3 * b, fW.x); atomicAdd(force + 3 * b + 1, fW.y); atomicAdd(force + 3 * b + 2, fW.z);
    atomicAdd(torque + 3 * b, tq.x); atomicAdd(torque + 3 * b + 1, tq.y); atomicAdd(torque + 3 * b + 2, tq.z);
  }
}

// Mouse tools (one thread): MouseSpringForce.apply (MouseSpringForce.java:69-101) and MouseImpulse.apply / Impulse
// (MouseImpulse.java:101-125, RigidBodySystem.java:249-267) as applyExternalForces runs them, before the scene's springs
struct MouseState {
  int springBody;  // -1: none
  int atCOM;
  int impBody;
  int impPhase;    // 0 none, 1 released (apply at the next applyExternalForces), 2 holding the force for one more application
  double grabB[3], pointW[3], k, c;
  double impPointB[3], impEndW[3], impScale;
  double heldPointW[3], heldForce[3];
};
__device__ __forceinline__ void wakeBody(int i, int* flags, int* metricCount);
__global__ void k_mouse_tools(int part /* 0: mouse spring (before the scene's springs), 1: impulse (after them) */, MouseState* M, const int* __restrict__ parent, int* __restrict__ flags, int* __restrict__ metricCount,
                              int* __restrict__ picked, const double* __restrict__ x, const double* __restrict__ R,
                              const double* __restrict__ v, const double* __restrict__ w, double* __restrict__ force,
                              double* __restrict__ torque) {
  if (blockIdx.x || threadIdx.x) return;
  if (part == 0 && M->springBody >= 0) {
    int b = M->springBody, p = parent[b];
    wakeBody(b, flags, metricCount);
    if (p >= 0) wakeBody(p, flags, metricCount);
    picked[b] = 1;
    xf T;
    T.R = ldm(R + 9 * b); T.t = ld3(x + 3 * b);
    d3 gW = xfP(T, ld3(M->grabB)), pW = ld3(M->pointW);
    d3 dd = vsub(gW, pW);
    double distance = sqrt(dd.x * dd.x + dd.y * dd.y + dd.z * dd.z);
    d3 dir = vsub(pW, gW);
    if (dir.x * dir.x + dir.y * dir.y + dir.z * dir.z >= 1e-3) {
      dir = vnormalize(dir);
      d3 at = M->atCOM ? ld3(x + 3 * b) : gW;
      d3 f = vscale(distance * M->k, dir);
      if (p >= 0) applyForceW(force, torque, x, p, gW, f, false);
      applyForceW(force, torque, x, b, at, f, false);
      d3 gv = spatialVelocity(x, v, w, b, gW);
      f = vscale(-vdot(gv, dir) * M->c, dir);
      if (p >= 0) applyForceW(force, torque, x, p, gW, f, false);
      applyForceW(force, torque, x, b, at, f, false);
    }
  }
  if (part == 0) return;
  if (M->impPhase == 1) {
    int b = M->impBody, p = parent[b];
    wakeBody(b, flags, metricCount);
    if (p >= 0) wakeBody(p, flags, metricCount);
    picked[b] = 1;
    xf T;
    T.R = ldm(R + 9 * b); T.t = ld3(x + 3 * b);
    d3 pW = xfP(T, ld3(M->impPointB)), eW = ld3(M->impEndW);
    d3 dd = vsub(eW, pW);
    double distance = sqrt(dd.x * dd.x + dd.y * dd.y + dd.z * dd.z);
    d3 dir = vsub(pW, eW);
    M->impPhase = 0;
    if (dir.x * dir.x + dir.y * dir.y + dir.z * dir.z >= 1e-3) {
      dir = vnormalize(dir);
      d3 f = vscale(M->impScale * distance, dir);
      if (p >= 0) applyForceW(force, torque, x, p, pW, f, false);
      applyForceW(force, torque, x, b, pW, f, false);
      st3(M->heldPointW, pW); st3(M->heldForce, f);
      M->impPhase = 2;
    }
  } else if (M->impPhase == 2) {  // RigidBodySystem.applyImpulse :262-267
    applyForceW(force, torque, x, M->impBody, ld3(M->heldPointW), ld3(M->heldForce), false);
    M->impPhase = 0;
  }
}

// One thread per body that has springs; its (spring, side) entries are visited in spring-list order so
// the accumulation order into force/torque is the reference's (applySpringForces :309-323).
__global__ void k_springs(int nsb, const int* __restrict__ spBodies, const int* __restrict__ start,
                          const int* __restrict__ list, const int* __restrict__ spType, const int* __restrict__ spB1,
                          const int* __restrict__ spB2, const double* __restrict__ pb1, const double* __restrict__ pb2,
                          const double* __restrict__ pw, const double* __restrict__ K, const double* __restrict__ D,
                          const double* __restrict__ L0, const double* __restrict__ LS, double ks, double ds,
                          const int* __restrict__ parent, const double* __restrict__ x, const double* __restrict__ R,
                          const double* __restrict__ v, const double* __restrict__ w, double* __restrict__ force,
                          double* __restrict__ torque) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsb) return;
  int body = spBodies[t];
  for (int e = start[t]; e < start[t + 1]; e++) {
    int s = list[e] >> 1, side = list[e] & 1;
    int b1 = spB1[s];
    xf T1;
    T1.R = ldm(R + 9 * b1);
    T1.t = ld3(x + 3 * b1);
    d3 p1W = xfP(T1, ld3(pb1 + 3 * s));
    d3 f;
    d3 at;
    if (spType[s] == AM3D_SPRING_BODYBODY) {
      int b2 = spB2[s];
      xf T2;
      T2.R = ldm(R + 9 * b2);
      T2.t = ld3(x + 3 * b2);
      d3 p2W = xfP(T2, ld3(pb2 + 3 * s));
      d3 disp = vsub(p2W, p1W);
      double dist = vlen(disp);
      if (dist < 1e-3) continue;
      disp = vscale(1. / dist, disp);
      d3 v2 = spatialVelocity(x, v, w, b2, p2W);
      d3 v1 = spatialVelocity(x, v, w, b1, p1W);
      d3 rel = vsub(v2, v1);
      double fi = K[s] * ks * (dist - L0[s] * LS[s]) + D[s] * ds * vdot(rel, disp);
      if (side == 0) { f = vscale(fi, disp); at = p1W; }
      else { f = vscale(-fi, disp); at = p2W; }
    } else {
      d3 disp = vsub(ld3(pw + 3 * s), p1W);
      double len = vlen(disp);
      if (len < 1e-3) continue;
      d3 v1 = spatialVelocity(x, v, w, b1, p1W);
      double sc = -(K[s] * ks * (len - L0[s] * LS[s]) - D[s] * ds * (vdot(v1, disp) / len)) / len;
      f = vscale(-sc, disp);
      at = p1W;
    }
    applyForceW(force, torque, x, body, at, f, false);
    if (parent[body] >= 0) applyForceW(force, torque, x, parent[body], at, f, true);
  }
}

// ------------------------------------------------------------------------------------------------
// body pairs
// ------------------------------------------------------------------------------------------------
__global__ void k_bpc_heads(int nc, const unsigned long long* __restrict__ key0, const int* __restrict__ b1,
                            const int* __restrict__ b2, const int* __restrict__ flags, int* __restrict__ head) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  unsigned long long k = key0[i] >> 16;
  bool skip = (flags[b1[i]] & AM3D_F_PINNED) && (flags[b2[i]] & AM3D_F_PINNED);  // storeInBodyPairContacts :173
  int h = 0;
  if (!skip) {
    if (i == 0) h = 1;
    else {
      unsigned long long kp = key0[i - 1] >> 16;
      h = (kp != k) ? 1 : 0;
    }
  }
  head[i] = h;
}
__global__ void k_bpc_fill(int nc, const unsigned long long* __restrict__ key0, const int* __restrict__ head,
                           const int* __restrict__ scan, const int* __restrict__ cb1, const int* __restrict__ cb2,
                           const int* __restrict__ flags, int* __restrict__ cbpc, unsigned long long* __restrict__ bkey,
                           int* __restrict__ bstart, int* __restrict__ bb1, int* __restrict__ bb2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  bool skip = (flags[cb1[i]] & AM3D_F_PINNED) && (flags[cb2[i]] & AM3D_F_PINNED);
  int b = scan[i] + head[i] - 1;
  cbpc[i] = skip ? -1 : b;
  if (head[i]) {
    bkey[b] = key0[i] >> 16;
    bstart[b] = i;
    bb1[b] = cb1[i];
    bb2[b] = cb2[i];
  }
}
// count + carry the persistent part (histories, creation orientation) over from last step's table
__global__ void k_bpc_match(int nb, int nc, const unsigned long long* __restrict__ key, const int* __restrict__ start,
                            int* __restrict__ count, int* __restrict__ b1, int* __restrict__ b2, int* __restrict__ nActive,
                            double* __restrict__ mh, int* __restrict__ sh, int* __restrict__ nm, int* __restrict__ nst,
                            int* __restrict__ alive, int np, const unsigned long long* __restrict__ pkey,
                            const int* __restrict__ pb1, const int* __restrict__ pb2, const double* __restrict__ pmh,
                            const int* __restrict__ psh, const int* __restrict__ pnm, const int* __restrict__ pnst,
                            int* __restrict__ maxCount) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  int end = (b + 1 < nb) ? start[b + 1] : nc;
  count[b] = end - start[b];
  if (end - start[b] > 64) atomicMax(maxCount, end - start[b]);  // sphere-tree pairs: next step's warm start indexes them
  nActive[b] = 0;
  alive[b] = 1;
  unsigned long long k = key[b];
  int lo = 0, hi = np;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (pkey[mid] < k) lo = mid + 1; else hi = mid;
  }
  if (lo < np && pkey[lo] == k) {
    b1[b] = pb1[lo];
    b2[b] = pb2[lo];
    nm[b] = pnm[lo];
    nst[b] = pnst[lo];
#pragma unroll
    for (int j = 0; j < 4; j++) { mh[4 * b + j] = pmh[4 * lo + j]; sh[4 * b + j] = psh[4 * lo + j]; }
  } else {
    nm[b] = 0;
    nst[b] = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { mh[4 * b + j] = 0; sh[4 * b + j] = 0; }
  }
}

// ------------------------------------------------------------------------------------------------
// warm start: one thread per body pair, contacts of the pair in emission order
// ------------------------------------------------------------------------------------------------
struct WarmCtx {
  // current
  const int *b1, *b2, *s1, *s2, *leaf;
  const unsigned long long *key0, *key1;
  const double* pB1;
  double *lam, *lamWarm, *prevViol;
  int* isNew;
  // previous (entries [0,npSorted) are in canonical key order, [npSorted,np) were appended by an unmerge)
  int np, npSorted, ncCur;
  const unsigned long long* tailKey;  // key0 of the appended entries, sorted
  const int* tailIdx;                 // their positions in the previous contact arrays
  const unsigned long long *pkey0, *pkey1;
  const int *pb1, *pleaf;
  const double *ppB1, *pviol;
  double* plam;
  // optional (key0,key1)-sorted index of the canonical entries (scenes with sphere trees: a pair can hold 10^4+ contacts)
  int useIdx;
  const unsigned long long *sk0, *sk1;
  const int* sidx;
  // bodies / shapes
  const int *btype, *shType;
  const double *x, *R;
  const int* ndRank;
  // box-box fast path: contact -> body pair, pairs that need the sequential repair, direct matches
  const int* bpc;
  int* pairSlow;
  int* matchIdx;
};

// find the previous contact equal to (key0, key1); duplicate keys (box x tree leaf hits, Appendix B):
// the reference's HashMap keeps the LAST one put, i.e. the last in DFS emission order = largest node rank
__device__ __forceinline__ int warmLookup(const WarmCtx& W, int lo, int hi, unsigned long long k0, unsigned long long k1) {
  int best = -1, bestRank = -1;
  if (W.useIdx) {
    int l = 0, h = W.npSorted;
    while (l < h) {
      int mid = (l + h) >> 1;
      unsigned long long a = W.sk0[mid];
      if (a < k0 || (a == k0 && W.sk1[mid] < k1)) l = mid + 1; else h = mid;
    }
    for (int t = l; t < W.npSorted && W.sk0[t] == k0 && W.sk1[t] == k1; t++) {  // equal keys stay in list order
      int j = W.sidx[t];
      int lf = W.pleaf[j];
      int rk = lf >= 0 ? W.ndRank[lf] : 0;
      if (best < 0 || rk > bestRank) { best = j; bestRank = rk; }
    }
  } else {
    for (int j = lo; j < hi; j++) {
      if (W.pkey1[j] == k1) {
        int lf = W.pleaf[j];
        int rk = lf >= 0 ? W.ndRank[lf] : 0;
        if (best < 0 || rk > bestRank) { best = j; bestRank = rk; }
      }
    }
  }
  int nt = W.np - W.npSorted;
  if (nt > 0) {  // entries appended later in the list (by an unmerge) win over earlier duplicates
    // (among themselves the duplicates of a box x tree pair again follow the reference's emission order: the contacts
    // of an unmerged pair return in their BodyPairContact.contactList order, i.e. by node rank - not in the order this
    // library happened to store them at the merge)
    int tl = 0, th = nt;
    while (tl < th) { int mid = (tl + th) >> 1; if (W.tailKey[mid] < k0) tl = mid + 1; else th = mid; }
    int tbest = -1, tRank = -1;
    for (int t = tl; t < nt && W.tailKey[t] == k0; t++) {
      int j = W.tailIdx[t];
      if (W.pkey1[j] == k1) {
        int lf = W.pleaf[j];
        int rk = lf >= 0 ? W.ndRank[lf] : 0;
        if (tbest < 0 || rk >= tRank) { tbest = j; tRank = rk; }
      }
    }
    if (tbest >= 0) best = tbest;
  }
  return best;
}
__device__ __forceinline__ void warmTake(const WarmCtx& W, int i, int j, bool zeroDonor) {
  W.isNew[i] = 0;
  for (int k = 0; k < 3; k++) {
    double l = W.plam[3 * j + k];
    W.lam[3 * i + k] = l;
    W.lamWarm[3 * i + k] = l;
    if (zeroDonor) W.plam[3 * j + k] = 0;
  }
  W.prevViol[i] = W.pviol[j];
}
__device__ __forceinline__ d3 worldPoint(const double* x, const double* R, int b, const d3& pB) {
  xf T;
  T.R = ldm(R + 9 * b);
  T.t = ld3(x + 3 * b);
  return xfP(T, pB);
}

// range of previous contacts with the same (bodyLo, bodyHi, partLo, partHi).  `guess` = where the range would start if
// last step's list were this step's (both are sorted by the same key): the lower bound is bracketed by galloping from
// there, a handful of probes in a settled scene instead of the ~24 of a bisection over 10^7 entries.
__device__ __forceinline__ void warmRange(const WarmCtx& W, unsigned long long k0, int guess, int& rlo, int& rhi) {
  rlo = rhi = 0;
  if (W.useIdx) return;
  int n = W.npSorted;
  int lo, hi;
  if (n == 0) return;
  guess = min(max(guess, 0), n - 1);
  if (W.pkey0[guess] < k0) {  // answer in (guess, n]
    int step = 1;
    lo = guess + 1;
    hi = lo;
    while (hi < n && W.pkey0[hi] < k0) { lo = hi + 1; hi = min(n, hi + step); step <<= 1; }
  } else {                    // answer in [0, guess]
    int step = 1;
    hi = guess;
    lo = hi;
    while (lo > 0 && W.pkey0[lo - 1] >= k0) { hi = lo - 1; lo = max(0, lo - 1 - step); step <<= 1; }
  }
  while (lo < hi) { int mid = (lo + hi) >> 1; if (W.pkey0[mid] < k0) lo = mid + 1; else hi = mid; }
  rlo = lo;
  // ranges are short here (pairs with more than 64 contacts go through the sorted index): walk to the end
  while (lo < n && W.pkey0[lo] == k0) lo++;
  rhi = lo;
}

// One thread per contact.
// Pairs without a box-box part (CollisionProcessor.java:479-495): plain key lookup, nothing is taken away from the donor,
// so the contacts of a pair are independent.
// Box-box pairs (:505-640): the common case is that every contact of the pair finds last step's contact of the same
// (key, info) within 0.05 - then no contact looks at another one's donor and the order inside the pair does not matter:
// the direct match is recorded here and applied by k_warm_apply.  A contact that misses marks its PAIR for the
// sequential repair path (k_warm_start), which then redoes that pair from scratch exactly as the reference does.
__global__ void k_warm_start_plain(int nc, WarmCtx W) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  int t1 = W.btype[W.b1[i]], t2 = W.btype[W.b2[i]];
  bool boxy = (t1 == AM3D_BODY_BOX || t1 == AM3D_BODY_COMPOSITE) && (t2 == AM3D_BODY_BOX || t2 == AM3D_BODY_COMPOSITE);
  unsigned long long k0 = W.key0[i], k1 = W.key1[i];
  int rlo, rhi;
  if (!boxy) {
    warmRange(W, k0, (int)((long long)i * W.npSorted / max(nc, 1)), rlo, rhi);
    int j = warmLookup(W, rlo, rhi, k0, k1);
    if (j >= 0) warmTake(W, i, j, false); else W.isNew[i] = 1;
    return;
  }
  int pair = W.bpc[i];
  if (pair < 0) return;
  W.matchIdx[i] = -1;
  bool nb1 = t1 == AM3D_BODY_COMPOSITE && W.shType[W.s1[i]] != AM3D_SHAPE_BOX;
  bool nb2 = t2 == AM3D_BODY_COMPOSITE && W.shType[W.s2[i]] != AM3D_SHAPE_BOX;
  if (nb1 || nb2) { W.pairSlow[pair] = 1; return; }
  warmRange(W, k0, (int)((long long)i * W.npSorted / max(nc, 1)), rlo, rhi);
  int j = warmLookup(W, rlo, rhi, k0, k1);
  if (j < 0) { W.pairSlow[pair] = 1; return; }
  d3 pNew = worldPoint(W.x, W.R, W.b1[i], ld3(W.pB1 + 3 * i));
  d3 pOld = worldPoint(W.x, W.R, W.pb1[j], ld3(W.ppB1 + 3 * j));
  if (vdist(pNew, pOld) < 0.05) W.matchIdx[i] = j;
  else W.pairSlow[pair] = 1;
}
__global__ void k_warm_apply(int nc, WarmCtx W) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  int t1 = W.btype[W.b1[i]], t2 = W.btype[W.b2[i]];
  bool boxy = (t1 == AM3D_BODY_BOX || t1 == AM3D_BODY_COMPOSITE) && (t2 == AM3D_BODY_BOX || t2 == AM3D_BODY_COMPOSITE);
  if (!boxy) return;
  int pair = W.bpc[i];
  if (pair < 0 || W.pairSlow[pair]) return;
  warmTake(W, i, W.matchIdx[i], true);
}

__global__ void k_gather_u64(int n, const int* __restrict__ idx, const unsigned long long* __restrict__ src,
                             unsigned long long* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

__global__ void k_warm_start(int nbp, const int* __restrict__ bstart, const int* __restrict__ bcount,
                             const int* __restrict__ bb1, const int* __restrict__ bb2, WarmCtx W) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbp) return;
  int s = bstart[b], e = s + bcount[b];
  int t1 = W.btype[bb1[b]], t2 = W.btype[bb2[b]];
  bool boxy = (t1 == AM3D_BODY_BOX || t1 == AM3D_BODY_COMPOSITE) && (t2 == AM3D_BODY_BOX || t2 == AM3D_BODY_COMPOSITE);
  if (!boxy) return;  // k_warm_start_plain
  if (!W.pairSlow[b]) return;  // every contact matched directly: k_warm_apply
  unsigned long long lastK0 = ~0ULL;
  int rlo = 0, rhi = 0;
  for (int i = s; i < e; i++) {
    unsigned long long k0 = W.key0[i], k1 = W.key1[i];
    if (k0 != lastK0) { warmRange(W, k0, (int)((long long)i * W.npSorted / max(W.ncCur, 1)), rlo, rhi); lastK0 = k0; }  // same (bodies, parts) as the previous contact: same range
    bool vanillaOnly = !boxy;
    bool doBox = boxy;
    if (boxy) {
      // composite parts that are not boxes (CollisionProcessor.java:497-503)
      bool nb1 = W.btype[W.b1[i]] == AM3D_BODY_COMPOSITE && W.shType[W.s1[i]] != AM3D_SHAPE_BOX;
      bool nb2 = W.btype[W.b2[i]] == AM3D_BODY_COMPOSITE && W.shType[W.s2[i]] != AM3D_SHAPE_BOX;
      if (nb1) { vanillaOnly = true; doBox = false; }
      else if (nb2) { vanillaOnly = true; doBox = true; }  // vanilla, then falls through into the box-box repair
    }
    if (vanillaOnly) {
      int j = warmLookup(W, rlo, rhi, k0, k1);
      if (j >= 0) warmTake(W, i, j, false); else W.isNew[i] = 1;
    }
    if (!doBox) continue;
    d3 pNew = worldPoint(W.x, W.R, W.b1[i], ld3(W.pB1 + 3 * i));
    int myInfo = (int)(k1 & 0xff);
    unsigned long long kbase = k1 & ~0xffULL;
    int j = warmLookup(W, rlo, rhi, k0, k1);
    if (j >= 0) {
      d3 pOld = worldPoint(W.x, W.R, W.pb1[j], ld3(W.ppB1 + 3 * j));
      double dist = vdist(pNew, pOld);
      if (dist > 0.05) {
        double bestDist = dist;
        int bestJ = j;
        for (int info = 0; info < 9; info++) {
          if (info == myInfo) continue;
          int jj = warmLookup(W, rlo, rhi, k0, kbase | (unsigned long long)info);
          if (jj < 0) break;
          pOld = worldPoint(W.x, W.R, W.pb1[jj], ld3(W.ppB1 + 3 * jj));
          dist = vdist(pNew, pOld);
          if (dist < bestDist) { bestDist = dist; bestJ = jj; }
        }
        j = bestJ;
        dist = bestDist;
      }
      if (dist < 0.05) warmTake(W, i, j, true); else W.isNew[i] = 1;
    } else {
      double bestDist = 1.7976931348623157e308;
      int bestJ = -1;
      for (int info = 0; info < 9; info++) {
        int jj = warmLookup(W, rlo, rhi, k0, kbase | (unsigned long long)info);
        if (jj < 0) break;
        d3 pOld = worldPoint(W.x, W.R, W.pb1[jj], ld3(W.ppB1 + 3 * jj));
        double dist = vdist(pNew, pOld);
        if (dist < bestDist) { bestDist = dist; bestJ = jj; }
      }
      if (bestJ >= 0) {
        if (bestDist < 0.05) warmTake(W, i, bestJ, true);
      } else {
        W.isNew[i] = 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// sleeping
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wakeBody(int b, int* flags, int* metricCount) {
  if (flags[b] & AM3D_F_SLEEPING) {
    atomicAnd(flags + b, ~AM3D_F_SLEEPING);
    metricCount[b] = 0;
  }
}
// Sleeping.wake :107-139, body-pair part
__global__ void k_wake_pairs(int nbp, const int* __restrict__ bb1, const int* __restrict__ bb2,
                             const int* __restrict__ parent, int* __restrict__ flags, int* __restrict__ metricCount) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbp) return;
  int b1 = bb1[b], b2 = bb2[b];
  if ((flags[b1] & AM3D_F_PINNED) || (flags[b2] & AM3D_F_PINNED)) return;
  int t1 = parent[b1] >= 0 ? parent[b1] : b1, t2 = parent[b2] >= 0 ? parent[b2] : b2;
  if ((flags[t1] & AM3D_F_SLEEPING) || (flags[t2] & AM3D_F_SLEEPING)) {
    wakeBody(b1, flags, metricCount); wakeBody(b2, flags, metricCount);
    wakeBody(t1, flags, metricCount); wakeBody(t2, flags, metricCount);
  }
}
__global__ void k_wake_springs(int nsp, const int* __restrict__ spType, const int* __restrict__ spB1,
                               const int* __restrict__ spB2, const int* __restrict__ parent, int* __restrict__ flags,
                               int* __restrict__ metricCount) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsp) return;
  if (spType[s] != AM3D_SPRING_BODYBODY) return;
  int b1 = spB1[s], b2 = spB2[s];
  int t1 = parent[b1] >= 0 ? parent[b1] : b1, t2 = parent[b2] >= 0 ? parent[b2] : b2;
  bool s1 = flags[t1] & AM3D_F_SLEEPING, s2 = flags[t2] & AM3D_F_SLEEPING;
  bool p1 = flags[t1] & AM3D_F_PINNED, p2 = flags[t2] & AM3D_F_PINNED;
  if (s1 != s2 && !(p1 || p2)) {
    wakeBody(b1, flags, metricCount); wakeBody(b2, flags, metricCount);
    wakeBody(t1, flags, metricCount); wakeBody(t2, flags, metricCount);
  }
}

__device__ __forceinline__ double bodyMetric(int b, const double* x, const double* R, const double* v, const double* w,
                                             const double* bbB, const int* bbCount) {
  double largest = 0;
  xf T;
  T.R = ldm(R + 9 * b);
  T.t = ld3(x + 3 * b);
  int n = bbCount[b];
  for (int k = 0; k < n; k++) {
    d3 pW = xfP(T, ld3(bbB + 24 * b + 3 * k));
    d3 v1 = spatialVelocity(x, v, w, b, pW);
    largest = fmax(vlen(v1), largest);
  }
  return largest;
}
// RigidBody.advancePositions :427-441 on a copy of the pose (expRodrigues :382-401, Matrix3d.normalizeCP)
__device__ __forceinline__ xf advancedPose(const xf& T, const d3& vv, const d3& om, double dt) {
  xf A;
  A.t = vscaleAdd(dt, vv, T.t);
  A.R = T.R;
  double t = vlen(om) * dt;
  if (t > 1e-8) {
    d3 wn = vnormalize(om);
    double c = cos(t), s = sin(t);
    double c1 = 1 - c;
    m3 dR;
    dR.m[0] = c + wn.x * wn.x * c1;
    dR.m[3] = wn.z * s + wn.x * wn.y * c1;
    dR.m[6] = -wn.y * s + wn.x * wn.z * c1;
    dR.m[1] = -wn.z * s + wn.x * wn.y * c1;
    dR.m[4] = c + wn.y * wn.y * c1;
    dR.m[7] = wn.x * s + wn.y * wn.z * c1;
    dR.m[2] = wn.y * s + wn.x * wn.z * c1;
    dR.m[5] = -wn.x * s + wn.y * wn.z * c1;
    dR.m[8] = c + wn.z * wn.z * c1;
    dR = mmul(dR, T.R);
    A.R = mnormalizeCP(dR);
  }
  return A;
}
// metricPositionLevel (MotionMetricProcessor.java:75-116): how far the bounding-box points of each body move in the
// other body's frame over one position update, divided by dt
__device__ __forceinline__ double pairMetricPos(int a, int b, const double* x, const double* R, const double* v, const double* w,
                                                const double* bbB, const int* bbCount, double dt) {
  xf T1, T2;
  T1.R = ldm(R + 9 * a); T1.t = ld3(x + 3 * a);
  T2.R = ldm(R + 9 * b); T2.t = ld3(x + 3 * b);
  xf A1 = advancedPose(T1, ld3(v + 3 * a), ld3(w + 3 * a), dt), A2 = advancedPose(T2, ld3(v + 3 * b), ld3(w + 3 * b), dt);
  double largest = 0, inv = 1. / dt;
  int n = bbCount[a];
  for (int k = 0; k < n; k++) {
    d3 point = ld3(bbB + 24 * a + 3 * k);
    d3 pB = xfInvP(T2, xfP(T1, point));
    pB = xfInvP(A1, xfP(A2, pB));
    largest = fmax(vlen(vscale(inv, vsub(point, pB))), largest);
  }
  n = bbCount[b];
  for (int k = 0; k < n; k++) {
    d3 point = ld3(bbB + 24 * b + 3 * k);
    d3 pB = xfInvP(T1, xfP(T2, point));
    pB = xfInvP(A2, xfP(A1, pB));
    largest = fmax(vlen(vscale(inv, vsub(point, pB))), largest);
  }
  return largest;
}
__device__ __forceinline__ double pairMetric(int a, int b, const int* flags, const double* x, const double* R,
                                             const double* v, const double* w, const double* bbB, const int* bbCount,
                                             double posDt = 0.0 /* > 0: metric at position level */) {
  if (flags[a] & AM3D_F_PINNED) return bodyMetric(b, x, R, v, w, bbB, bbCount);
  if (flags[b] & AM3D_F_PINNED) return bodyMetric(a, x, R, v, w, bbB, bbCount);
  if (posDt > 0.0) return pairMetricPos(a, b, x, R, v, w, bbB, bbCount, posDt);
  double largest = 0;
  for (int i = 0; i < 2; i++) {
    int body = i == 0 ? a : b;
    xf T;
    T.R = ldm(R + 9 * body);
    T.t = ld3(x + 3 * body);
    int n = bbCount[body];
    for (int k = 0; k < n; k++) {
      d3 pW = xfP(T, ld3(bbB + 24 * body + 3 * k));
      d3 v1 = spatialVelocity(x, v, w, a, pW);
      d3 v2 = spatialVelocity(x, v, w, b, pW);
      v1 = vsub(v1, v2);
      largest = fmax(vlen(v1), largest);
    }
  }
  return largest;
}

// Sleeping.sleep :49-98 for top-level bodies
__global__ void k_sleep(int ns, int nb, const int* __restrict__ alive, const int* __restrict__ parent,
                        int* __restrict__ flags, const int* __restrict__ hasExt, double* __restrict__ hist,
                        int* __restrict__ hcount, const double* __restrict__ x, const double* __restrict__ R,
                        const double* __restrict__ v, const double* __restrict__ w, const double* __restrict__ bbB,
                        const int* __restrict__ bbCount, int stepAccum, double threshold) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  if (i >= nb ? !alive[i - nb] : parent[i] >= 0) return;
  if (flags[i] & AM3D_F_SLEEPING) return;
  if (hasExt[i]) return;
  double m = bodyMetric(i, x, R, v, w, bbB, bbCount);
  int n = hcount[i];
  double* h = hist + 10 * i;
  if (stepAccum > 10) stepAccum = 10;
  // metricHistory.add(m); if (size > stepAccum) remove(0)   (Sleeping.java:153-158)
  if (n < 10) h[n++] = m;
  else { for (int k = 0; k + 1 < 10; k++) h[k] = h[k + 1]; h[9] = m; }
  if (n > stepAccum) { for (int k = 0; k + 1 < n; k++) h[k] = h[k + 1]; n--; }
  hcount[i] = n;
  bool sleep = true;
  double prev = 1.7976931348623157e308;
  if (n < stepAccum) {
    sleep = false;
  } else {
    for (int k = 0; k < n; k++) {
      double mk = h[k];
      if (mk > prev + 5e-5) { sleep = false; break; }
      if (mk > threshold) { sleep = false; break; }
      prev = mk;
    }
  }
  if (sleep) flags[i] |= AM3D_F_SLEEPING;
}

// ------------------------------------------------------------------------------------------------
// graph colouring of the body-pair groups (Jones-Plassmann with per-body colour masks)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int hash32(unsigned int a) {
  a = (a ^ 61) ^ (a >> 16);
  a *= 9;
  a = a ^ (a >> 4);
  a *= 0x27d4eb2d;
  a = a ^ (a >> 15);
  return a;
}
// solver bodies of a group: parent collection unless computeInCollection; negative (-1 - body) for a pinned body
__global__ void k_grp_init(int ng, const int* __restrict__ gb1, const int* __restrict__ gb2, const int* __restrict__ parent,
                           const int* __restrict__ flags, const int* __restrict__ gcount, const int* __restrict__ bodyLocal,
                           const int* __restrict__ lead /* chunked pairs: 1 on the first chunk of a pair; or nullptr */, int inCollection,
                           int* __restrict__ sb1, int* __restrict__ sb2, unsigned long long* __restrict__ prio,
                           int* __restrict__ color, int* __restrict__ degree) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  int a = gb1[g], b = gb2[g];
  if (!inCollection) {
    if (parent[a] >= 0) a = parent[a];
    if (parent[b] >= 0) b = parent[b];
  }
  int cnt = gcount[g];
  // groups without contacts in this solve (dead pairs, internal pairs of sleeping collections the sweep does not reach)
  // constrain nobody
  bool pa = (flags[a] & AM3D_F_PINNED) || cnt == 0, pb = (flags[b] & AM3D_F_PINNED) || cnt == 0;
  sb1[g] = pa ? -1 - a : a;
  sb2[g] = pb ? -1 - b : b;
  if (!lead || lead[g]) {  // degree in body pairs (decides what is a hub), not in chunks
    if (!pa) atomicAdd(degree + a, 1);
    if (!pb) atomicAdd(degree + b, 1);
  }
  // Jones-Plassmann priority: hashed, except that long groups (sphere-tree pairs with hundreds of contacts, solved
  // as one sequential chain) come first by size class: giants that do not touch each other then share the first
  // colours, and a sweep costs the longest chain per colour instead of one giant per colour
  unsigned long long cls = cnt > 64 ? (unsigned long long)(32 - __clz(cnt >> 6)) : 0ULL;
  // the hash is taken over the SCENE-LOCAL ids of the two leaf bodies and ties go to the lower group index, whose
  // order inside a scene does not depend on the other scenes of the context: a scene is coloured (and therefore
  // solved) the same way alone and as one of many batched copies
  unsigned la = (unsigned)bodyLocal[gb1[g]], lb = (unsigned)bodyLocal[gb2[g]];
  unsigned h = hash32(hash32(la) * 0x9E3779B1u + lb);
  prio[g] = (cls << 59) | ((unsigned long long)(h >> 5) << 32) | (unsigned)(g + 1);
  color[g] = -1;
}
// Hubs: non-pinned solver bodies touched by >= hubMin groups (a funnel, a big merged collection).  A hub side takes
// no part in the colouring (like a pinned body); its deltaV is updated once per colour from the per-group deltas.
__global__ void k_grp_hubs(int ng, const int* __restrict__ sb1, const int* __restrict__ sb2, const int* __restrict__ degree,
                           int hubMin, int* __restrict__ hubMask, int* __restrict__ nHubSides) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  int m = 0;
  if (hubMin > 0) {
    if (sb1[g] >= 0 && degree[sb1[g]] >= hubMin) m |= 1;
    if (sb2[g] >= 0 && degree[sb2[g]] >= hubMin) m |= 2;
  }
  hubMask[g] = m;
  if (m) atomicAdd(nHubSides, (m & 1) + ((m >> 1) & 1));
}
__global__ void k_color_bid(int ng, const int* __restrict__ sb1, const int* __restrict__ sb2, const int* __restrict__ hubMask,
                            const unsigned long long* __restrict__ prio, const int* __restrict__ color,
                            unsigned long long* __restrict__ best) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng || color[g] != -1) return;
  unsigned long long p = prio[g];
  int hm = hubMask[g];
  if (sb1[g] >= 0 && !(hm & 1)) atomicMax(best + sb1[g], p);
  if (sb2[g] >= 0 && !(hm & 2)) atomicMax(best + sb2[g], p);
}
// page = which block of 64 colours is being filled; groups whose bodies have no free colour left in this
// page are deferred to the next one (color = -2 - page marks "waiting for page+1")
__global__ void k_color_assign(int ng, int page, const int* __restrict__ sb1, const int* __restrict__ sb2,
                               const int* __restrict__ hubMask, const unsigned long long* __restrict__ prio, int* __restrict__ color,
                               unsigned long long* __restrict__ best, unsigned long long* __restrict__ mask,
                               int* __restrict__ remaining, int* __restrict__ deferred) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng || color[g] != -1) return;
  unsigned long long p = prio[g];
  int a = sb1[g], b = sb2[g];
  int hm = hubMask[g];
  if (hm & 1) a = -1;  // hub sides do not constrain the colour
  if (hm & 2) b = -1;
  bool win = (a < 0 || best[a] == p) && (b < 0 || best[b] == p);
  if (!win) { atomicAdd(remaining, 1); return; }
  unsigned long long m = (a >= 0 ? mask[a] : 0ULL) | (b >= 0 ? mask[b] : 0ULL);
  if (~m == 0ULL) {
    color[g] = -2 - page;
    atomicAdd(deferred, 1);
  } else {
    int c = __ffsll((long long)~m) - 1;
    color[g] = page * 64 + c;
    if (a >= 0) mask[a] |= 1ULL << c;
    if (b >= 0) mask[b] |= 1ULL << c;
  }
  if (a >= 0) best[a] = 0;
  if (b >= 0) best[b] = 0;
}
__global__ void k_color_next_page(int ng, int page, int* __restrict__ color) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  if (color[g] == -2 - page) color[g] = -1;
}
// The whole Jones-Plassmann colouring in ONE cooperative launch: the bid / assign rounds of k_color_bid and k_color_assign
// with grid barriers between them, pages of 64 colours until no group is deferred.  ctl: [0..2] "a round left groups
// uncoloured" flags (rotating), [3] deferred groups of the current page, [4] pages used (out), [5] error: too many pages
__global__ void __launch_bounds__(256)
k_color_coop(int ng, int ns, int maxPages, const int* __restrict__ sb1, const int* __restrict__ sb2, const int* __restrict__ hubMask,
             const unsigned long long* __restrict__ prio, int* __restrict__ color, unsigned long long* __restrict__ best,
             unsigned long long* __restrict__ mask, int* __restrict__ ctl) {
  cg::grid_group grid = cg::this_grid();
  int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  int round = 0;
  for (int page = 0;; page++) {
    for (;;) {
      round++;
      if (tid == 0) ctl[(round + 1) % 3] = 0;
      for (int g = tid; g < ng; g += stride) {
        if (color[g] != -1) continue;
        unsigned long long p = prio[g];
        int hm = hubMask[g];
        if (sb1[g] >= 0 && !(hm & 1)) atomicMax(best + sb1[g], p);
        if (sb2[g] >= 0 && !(hm & 2)) atomicMax(best + sb2[g], p);
      }
      grid.sync();
      bool left = false;
      for (int g = tid; g < ng; g += stride) {
        if (color[g] != -1) continue;
        unsigned long long p = prio[g];
        int a = sb1[g], b = sb2[g];
        int hm = hubMask[g];
        if (hm & 1) a = -1;
        if (hm & 2) b = -1;
        bool win = (a < 0 || __ldcg(best + a) == p) && (b < 0 || __ldcg(best + b) == p);
        if (!win) { left = true; continue; }
        unsigned long long m = (a >= 0 ? __ldcg(mask + a) : 0ULL) | (b >= 0 ? __ldcg(mask + b) : 0ULL);
        if (~m == 0ULL) {
          color[g] = -2 - page;
          atomicAdd(ctl + 3, 1);
        } else {
          int c = __ffsll((long long)~m) - 1;
          color[g] = page * 64 + c;
          if (a >= 0) mask[a] = __ldcg(mask + a) | (1ULL << c);
          if (b >= 0) mask[b] = __ldcg(mask + b) | (1ULL << c);
        }
        if (a >= 0) best[a] = 0;
        if (b >= 0) best[b] = 0;
      }
      if (left) ctl[round % 3] = 1;
      grid.sync();
      if (__ldcg(ctl + round % 3) == 0) break;
    }
    int deferred = __ldcg(ctl + 3);
    if (deferred == 0) { if (tid == 0) ctl[4] = page + 1; break; }
    if (page + 1 >= maxPages) { if (tid == 0) { ctl[5] = 1; ctl[4] = page + 1; } break; }
    grid.sync();  // everybody has read the deferred count
    if (tid == 0) ctl[3] = 0;
    for (int i = tid; i < ns; i += stride) mask[i] = 0ULL;
    for (int g = tid; g < ng; g += stride) if (color[g] == -2 - page) color[g] = -1;
    grid.sync();
  }
}

// solve order: by (layer, colour) = one PHASE of the sweep; inside a phase (any order gives the same result: the groups
// share no free body) by descending contact count so that the lanes of a warp run the same number of contacts; ties by
// group index (the radix sort is stable).  layer = breadth-first distance from the body pairs that hold new contacts
// (getOrganizedContacts, CollisionProcessor.java:346-441) for the single sweep, absent (0) for the full solve.
// partShift > 0: batched scenes are solved in PARTITIONS (blocks of consecutive scenes, one thread-block cluster each):
// the partition leads the key, so that every partition's phases are contiguous
__global__ void k_color_sortkey(int ng, const int* __restrict__ color, const int* __restrict__ gcount, const int* __restrict__ layer,
                                const int* __restrict__ gb1, const int* __restrict__ bodyScene, int partShift, int nPart, int nScenes,
                                unsigned long long* __restrict__ key, int* __restrict__ val) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  unsigned long long L = layer ? (unsigned long long)(unsigned)layer[g] : 0ULL;
  unsigned long long k = (L << 20) | ((unsigned long long)(unsigned)color[g] << 8) | (unsigned)(255 - min(gcount[g], 255));
  if (partShift > 0) k |= (unsigned long long)((long long)bodyScene[gb1[g]] * nPart / nScenes) << partShift;
  key[g] = k;
  val[g] = g;
}
// per partition: its range of phases (partitions without groups keep begin = end = 0)
__global__ void k_part_phases(int nPhases, const int* __restrict__ phaseStart, const unsigned long long* __restrict__ key, int partShift,
                              int* __restrict__ partRange /* [2 * nPart] */) {
  int ph = blockIdx.x * blockDim.x + threadIdx.x;
  if (ph >= nPhases) return;
  int s = (int)(key[phaseStart[ph]] >> partShift);
  if (ph == 0 || (int)(key[phaseStart[ph - 1]] >> partShift) != s) partRange[2 * s] = ph;
  if (ph == nPhases - 1 || (int)(key[phaseStart[ph + 1]] >> partShift) != s) partRange[2 * s + 1] = ph + 1;
}
// phases of the sorted group list: head = first group of a (layer, colour) class
__global__ void k_phase_heads(int ng, const unsigned long long* __restrict__ key, int* __restrict__ head) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ng) return;
  head[p] = (p == 0 || (key[p] >> 8) != (key[p - 1] >> 8)) ? 1 : 0;
}
__global__ void k_phase_fill(int ng, const int* __restrict__ head, const int* __restrict__ scan, int* __restrict__ phaseStart,
                             int* __restrict__ phaseOf) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ng) return;
  int ph = scan[p] + head[p] - 1;
  phaseOf[p] = ph;
  if (head[p]) phaseStart[ph] = p;
  if (p == ng - 1) phaseStart[ph + 1] = ng;
}

// ------------------------------------------------------------------------------------------------
// solve-order setup + assembly
// ------------------------------------------------------------------------------------------------
struct SolveArrays {
  int *sgB1, *sgB2, *sgStart, *sgCount, *sgFlags, *sgBpc;
  int* sgScene;      // scene of the group (a context can hold many independent scenes)
  int* sceneState;   // done[nScenes] | iterations executed[nScenes] | still moving in this iteration[nScenes][MV_SLOTS]  (PGS.java:190-192 per scene)
  double *sgMass, *sgMu;
  double* scP;       // [24] per contact, solve order: n t1 t2 (9) | r1 r2 (6) | b (3) | D (3) | lambda (3)
  int *scSrc, *scState;
  double* hubDelta;  // [12] per group: what this group added to a hub body on side 1 / side 2 in the current colour
};
#define SG_CLAMP 1
#define SG_HUB1 2
#define SG_HUB2 4

// one thread per group in solve order: gather body data (mass packet, friction, magnet flags)
__global__ void k_group_setup(int ng, const int* __restrict__ order /* solve pos -> group */, const int* __restrict__ sb1,
                              const int* __restrict__ sb2, const int* __restrict__ gb1, const int* __restrict__ gb2,
                              const int* __restrict__ gcount, const double* __restrict__ minv,
                              const double* __restrict__ jinv, const double* __restrict__ fric,
                              const int* __restrict__ flags, const int* __restrict__ hubMask, int frictionOverride,
                              double frictionVal, const int* __restrict__ bodyScene, SolveArrays S, int* __restrict__ grpPos) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ng) return;
  int g = order[p];
  grpPos[g] = p;
  int a = sb1[g], b = sb2[g];
  int ia = a >= 0 ? a : -1 - a, ib = b >= 0 ? b : -1 - b;
  S.sgB1[p] = a >= 0 ? a : -1;
  S.sgB2[p] = b >= 0 ? b : -1;
  S.sgScene[p] = bodyScene[gb1[g]];
  S.sgCount[p] = gcount[g];
  S.sgBpc[p] = g;
  double* M = S.sgMass + 20 * p;
  M[0] = minv[ia];
  for (int k = 0; k < 9; k++) M[1 + k] = jinv[9 * ia + k];
  M[10] = minv[ib];
  for (int k = 0; k < 9; k++) M[11 + k] = jinv[9 * ib + k];
  int l1 = gb1[g], l2 = gb2[g];
  double mu;
  if (frictionOverride) mu = frictionVal;
  else {
    double f1 = fric[l1], f2 = fric[l2];
    if (f1 < 0.2 || f2 < 0.2) mu = fmin(f1, f2);
    else if (f1 > 1. || f2 > 1.) mu = fmax(f1, f2);
    else mu = (f1 + f2) / 2.;
  }
  S.sgMu[p] = mu;
  int fl1 = flags[l1], fl2 = flags[l2];
  bool clamp = (!(fl1 & AM3D_F_MAGNETIC) || !(fl1 & AM3D_F_MAGNET_ACTIVE)) && (!(fl2 & AM3D_F_MAGNETIC) || !(fl2 & AM3D_F_MAGNET_ACTIVE));
  S.sgFlags[p] = (clamp ? SG_CLAMP : 0) | ((hubMask[g] & 1) ? SG_HUB1 : 0) | ((hubMask[g] & 2) ? SG_HUB2 : 0);
}

// Contact.computeJacobian :235-271, computeB :279-326, computeJMinvJt :340-354; one thread per contact,
// output in solve order.  Body state is read through the solver-body index (collection parent unless
// inCollection).
__global__ void k_assemble(int nc, const int* __restrict__ cbpc, int groupOffset, int setId,
                           const int* __restrict__ gstart /* body pair -> first contact in its set */,
                           const int* __restrict__ gcount, const int* __restrict__ chunkFirst /* pair -> first chunk, or nullptr */,
                           int chunkLen, const int* __restrict__ chunkStart, const int* __restrict__ grpPos,
                           const int* __restrict__ sgStart, const int* __restrict__ cb1,
                           const int* __restrict__ cb2, const int* __restrict__ parent, int inCollection,
                           const double* __restrict__ pW, const double* __restrict__ nW, const double* __restrict__ t1W,
                           const double* __restrict__ t2W, const double* __restrict__ pB1, const double* __restrict__ nB1,
                           const double* __restrict__ t1B1, const double* __restrict__ t2B1, const double* __restrict__ viol,
                           const double* __restrict__ lam, const int* __restrict__ cstate, const double* __restrict__ x,
                           const double* __restrict__ R, const double* __restrict__ v, const double* __restrict__ w,
                           const double* __restrict__ force, const double* __restrict__ torque,
                           const double* __restrict__ minv, const double* __restrict__ jinv, const double* __restrict__ rest,
                           double dt, double feedback, int postStab, int restOverride, double restVal, SolveArrays S) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  int g = cbpc[i];
  if (g < 0) return;
  g += groupOffset;
  if (gcount[g] == 0) return;  // internal pair of a sleeping collection: not part of the sweep
  int idx;
  if (chunkFirst) {
    int e = chunkFirst[g] + (chunkFirst[g + 1] - chunkFirst[g] > 1 ? (i - gstart[g]) / chunkLen : 0);
    idx = sgStart[grpPos[e]] + (i - chunkStart[e]);
  } else {
    idx = sgStart[grpPos[g]] + (i - gstart[g]);
  }
  int l1 = cb1[i], l2 = cb2[i];
  int a = l1, b = l2;
  bool touchesCollection = parent[l1] >= 0 || parent[l2] >= 0;
  if (!inCollection) {
    if (parent[a] >= 0) a = parent[a];
    if (parent[b] >= 0) b = parent[b];
  }
  d3 p, n, t1, t2;
  if (touchesCollection) {
    // updateJacobiansThatNeedUpdating :326-334: frame re-expressed from body1 coordinates
    xf T;
    T.R = ldm(R + 9 * l1);
    T.t = ld3(x + 3 * l1);
    p = xfP(T, ld3(pB1 + 3 * i));
    n = mtransform(T.R, ld3(nB1 + 3 * i));
    t1 = mtransform(T.R, ld3(t1B1 + 3 * i));
    t2 = mtransform(T.R, ld3(t2B1 + 3 * i));
  } else {
    p = ld3(pW + 3 * i); n = ld3(nW + 3 * i); t1 = ld3(t1W + 3 * i); t2 = ld3(t2W + 3 * i);
  }
  d3 r1 = vsub(p, ld3(x + 3 * a)), r2 = vsub(p, ld3(x + 3 * b));
  double* PKg = S.scP + 24 * (size_t)idx;
  double PK[24];
  st3(PK, n); st3(PK + 3, t1); st3(PK + 6, t2);
  st3(PK + 9, r1); st3(PK + 12, r2);
  // Jacobian rows
  d3 dir[3] = {n, t1, t2};
  d3 jaw[3], jbw[3];
#pragma unroll
  for (int k = 0; k < 3; k++) { jaw[k] = vcross(dir[k], r1); jbw[k] = vcross(r2, dir[k]); }
  double e = (rest[l1] + rest[l2]) / 2.;
  if (restOverride) e = restVal;
  double bb[3] = {0, 0, 0};
  d3 v1 = ld3(v + 3 * a), w1 = ld3(w + 3 * a), v2 = ld3(v + 3 * b), w2 = ld3(w + 3 * b);
  m3 J1 = ldm(jinv + 9 * a), J2 = ldm(jinv + 9 * b);
  double mi1 = minv[a], mi2 = minv[b];
  if (postStab) {  // PGS.java:86-89: position-level right-hand side
    bb[0] = feedback * viol[i];
  } else {
    d3 tmp = vscaleAdd(mi1 * dt, ld3(force + 3 * a), v1);
#pragma unroll
    for (int k = 0; k < 3; k++) bb[k] += vdot(tmp, vscale(-1, dir[k]));
    tmp = mtransform(J1, ld3(torque + 3 * a));
    tmp = vscale(dt, tmp);
    tmp = vadd(tmp, w1);
#pragma unroll
    for (int k = 0; k < 3; k++) bb[k] += vdot(tmp, jaw[k]);
    double bBounce = vdot(v1, vscale(-1, n)) + vdot(w1, jaw[0]);
    bBounce *= e;
    bb[0] += bBounce;
    tmp = vscaleAdd(mi2 * dt, ld3(force + 3 * b), v2);
#pragma unroll
    for (int k = 0; k < 3; k++) bb[k] += vdot(tmp, dir[k]);
    tmp = mtransform(J2, ld3(torque + 3 * b));
    tmp = vscale(dt, tmp);
    tmp = vadd(tmp, w2);
#pragma unroll
    for (int k = 0; k < 3; k++) bb[k] += vdot(tmp, jbw[k]);
    bBounce = vdot(v2, n) + vdot(w2, jbw[0]);
    bBounce *= e;
    bb[0] += bBounce;
    bb[0] += feedback * viol[i];
  }
  PK[15] = bb[0]; PK[16] = bb[1]; PK[17] = bb[2];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    d3 jav = vscale(-1, dir[k]);
    d3 tmp1 = mtransform(J1, jaw[k]), tmp2 = mtransform(J2, jbw[k]);
    PK[18 + k] = mi1 * vdot(jav, jav) + vdot(jaw[k], tmp1) + mi2 * vdot(dir[k], dir[k]) + vdot(jbw[k], tmp2);
    PK[21 + k] = lam[3 * i + k];
  }
#pragma unroll
  for (int k = 0; k < 6; k++) st4(PKg + 4 * k, PK[4 * k], PK[4 * k + 1], PK[4 * k + 2], PK[4 * k + 3]);
  S.scSrc[idx] = i | (setId << 30);
  S.scState[idx] = cstate[i];
}

// ------------------------------------------------------------------------------------------------
// PGS sweeps: one launch per colour, one thread per body-pair group, the group's contacts in sequence
// ------------------------------------------------------------------------------------------------
#define MV_SLOTS 32  // "still moving" flags per scene, spread over CTAs so that a single-scene context does not hammer one address
struct PgsParams {
  double omega, compliance, tolerance, sliding;
  int nScenes;  // sceneState = done[nScenes] | iterations[nScenes] | moving[nScenes][MV_SLOTS]
  int check;    // take the tolerance exit (full solve) or not (single sweep)
  int fastRows; // branch-free row update where it applies (option "pgs_fast_rows", default on; results are identical)
};

__device__ __forceinline__ double dot6(const d3& jv, const d3& jw, const double* dvp) {
  return (jv.x * dvp[0] + jv.y * dvp[1] + jv.z * dvp[2]) + (jw.x * dvp[3] + jw.y * dvp[4] + jw.z * dvp[5]);
}
__device__ __forceinline__ void applyRow(double* dvp, double minv, const double* J, const d3& jv, const d3& jw, double lambda) {
  double s = minv * lambda;
  dvp[0] = s * jv.x + dvp[0];
  dvp[1] = s * jv.y + dvp[1];
  dvp[2] = s * jv.z + dvp[2];
  double tx = J[0] * jw.x + J[1] * jw.y + J[2] * jw.z;
  double ty = J[3] * jw.x + J[4] * jw.y + J[5] * jw.z;
  double tz = J[6] * jw.x + J[7] * jw.y + J[8] * jw.z;
  dvp[3] = lambda * tx + dvp[3];
  dvp[4] = lambda * ty + dvp[4];
  dvp[5] = lambda * tz + dvp[5];
}

// One body-pair group: its contacts in sequence, the deltaV of its two solver bodies held in registers.
// MODE 0: confidentWarmStart (PGS.java:250-256); MODE 1: one Gauss-Seidel sweep (:105-181).
// A hub side works on a private copy of the hub's deltaV as of the start of the colour (plus this group's own
// updates) and hands what it added to hubDelta; k_hub_reduce folds the deltas in after the colour, in a fixed order.
__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// a / b, correctly rounded, from y = RN(1/b): q = a*y is within an ulp or two of the quotient, and each step
// q <- fma(fma(-b, q, a), y, q) keeps or improves it; once q is a faithful rounding the step returns RN(a/b) exactly
// (Markstein's theorem, the exact residual comes from the fused multiply-add).  600 M random and adversarial operand
// pairs against the IEEE division: no mismatch already after ONE step; two are taken.  Zero numerators keep their
// sign through a*y; results outside the normal range (underflow, overflow, NaN, y not finite) take the IEEE division.
// Why: the division of the row update sits on the dependent chain of a group, where the IEEE sequence (reciprocal
// seed + Newton steps + fix-up) costs ~30 dependent instructions; 1/b does not depend on the chain.
__device__ __forceinline__ double divExact(double a, double b, double y) {
  double q = a * y;
  if (a == 0.0) return q;
  double r = __fma_rn(-b, q, a);
  q = __fma_rn(r, y, q);
  r = __fma_rn(-b, q, a);
  q = __fma_rn(r, y, q);
  double m = fabs(q);
  if (!(m > 1e-280 && m < 1e280)) q = a / b;
  return q;
}
// --- branch-free forms of the row update, for the latency-bound chain ---
// Measured on B200 (tools/micro/fp64_latency.cu): a dependent DADD / DMUL / DFMA costs ~8.2 cycles, but fmax / fmin on
// doubles compile to DSETP + selects behind a NaN branch (~26 cycles each), and every branch on a DSETP result (the zero
// and range tests of divExact) stalls the in-order warp for the compare's latency.  The chain of a giant pair has nothing
// to overlap with, so its row update is written without any compare-and-branch: the quotient is the Markstein sequence
// unconditionally, the clamps are integer selects on the bit patterns, and the conditions under which those forms are not
// the reference's arithmetic (quotient outside [2^-929, 2^930] - this includes 0, denormals, Inf, NaN - or a friction
// bound that is not a non-negative finite number) only set a flag OFF the chain; a batch of 32 contacts that raised the
// flag is run again from its saved deltaV with the exact (branching) forms.  Results are bit-identical either way.
__device__ __forceinline__ double quotFast(double a, double b, double y, unsigned& bad) {
  double q0 = a * y;
  double r = __fma_rn(-b, q0, a);
  double q = __fma_rn(r, y, q0);
  r = __fma_rn(-b, q, a);
  q = __fma_rn(r, y, q);
  // a == +-0: the quotient is a * y (a signed zero); the fused steps would lose the sign of -0
  int ah = __double2hiint(a), al = __double2loint(a);
  bool azero = (((unsigned)ah << 1) | (unsigned)al) == 0u;
  unsigned e = ((unsigned)__double2hiint(q) >> 20) & 0x7ffu;
  bad |= (!azero && (e - 94u > 1859u)) ? 1u : 0u;
  return __hiloint2double(azero ? __double2hiint(q0) : __double2hiint(q), azero ? __double2loint(q0) : __double2loint(q));
}
// std::max(0.0, l) for a non-NaN l (PGS.java:120; -0.0 gives +0.0 as Math.max does)
__device__ __forceinline__ double clampNonNegBits(double l) {
  int hi = __double2hiint(l), lo = __double2loint(l);
  bool neg = hi < 0;
  return __hiloint2double(neg ? 0 : hi, neg ? 0 : lo);
}
// std::min(std::max(l, -limit), limit) for a finite limit >= +0 and a non-NaN l (PGS.java:151-154, 168-171): non-negative
// doubles order like their bit patterns
__device__ __forceinline__ double clampAbsBits(double l, double limit, unsigned& bad) {
  unsigned long long ul = (unsigned long long)__double_as_longlong(limit), uq = (unsigned long long)__double_as_longlong(l);
  bad |= (ul >= 0x7ff0000000000000ULL) ? 1u : 0u;  // negative, -0, Inf or NaN bound: take the exact path
  bool gt = (uq & 0x7fffffffffffffffULL) > ul;
  return __longlong_as_double((long long)(gt ? (ul | (uq & 0x8000000000000000ULL)) : uq));
}

// The tolerance exit of PGS.java:190-192 is taken PER SCENE: a context can hold many independent scenes (batched
// copies), and each of them leaves the iteration when ITS largest |delta lambda| falls below the tolerance, exactly as
// if it were solved alone.  A group whose |delta lambda| stays at or above the tolerance marks its scene as moving;
// scenes nobody marked are done after the iteration (k_iter_end / the persistent kernel) and their groups are skipped.
// (Measured and not kept: two lanes per group, one per body, meeting through a shuffle per row - halves the serial
// instruction stream of a group but doubles the load instructions and the redundant lambda arithmetic: slower on
// both the batched scenes and the 1M-box stack; staging a scene's records through shared memory with bulk copies
// (cp.async.bulk + mbarrier, 2 x 90 KB stages, one CTA per SM) removes the memory stalls but leaves one scene per
// SM: 3.5x slower than two register-heavy CTAs per SM.  ncu: the sweep is bound by the dependent FP64 instruction
// stream of a group, ~600 instructions per contact at one issue per ~5.6 cycles.  One CTA per scene with the scene's
// deltaV in shared memory and __syncthreads() for grid barriers: 1.4x slower at 512 scenes per GPU, 2.2x slower at
// 4096 than the grid-wide sweeps, which keep every SM busy with whatever scene has work.)
template <int MODE, bool HUB>
__device__ __forceinline__ void pgsGroup(int p, const SolveArrays& S, double* __restrict__ dv, const PgsParams& P, int lastIter) {
  // (the header loads are issued before the scene's done flag is looked at: the two dependent loads of that test would
  // otherwise delay everything behind them by a memory round trip per phase)
  int scene = (MODE == 1 && P.check) ? S.sgScene[p] : 0;
  double localMax = 0;
  int a = S.sgB1[p], b = S.sgB2[p];
  int start = S.sgStart[p], cnt = S.sgCount[p];
  const double* PK0 = S.scP + 24 * (size_t)start;
  if (cnt > 0) { prefetchL2(PK0); prefetchL2(PK0 + 16); }
  if (MODE == 1 && P.check && S.sceneState[scene]) return;  // this scene has left the iteration
  double M[20];
  const double* Mp = S.sgMass + 20 * (size_t)p;
#pragma unroll
  for (int k = 0; k < 5; k++) ld4(Mp + 4 * k, M + 4 * k);
  double mu = S.sgMu[p];
  int fl = S.sgFlags[p];
  bool clamp = fl & SG_CLAMP;
  const bool hubA = HUB && (fl & SG_HUB1), hubB = HUB && (fl & SG_HUB2);  // HUB = false: the solve has no hub body
  // branch-free rows: for clamped groups (no active magnet); a pinned side then carries a zeroed mass packet so that the
  // unguarded update leaves its (never stored) deltaV at +0
  const bool useFast = MODE == 1 && P.fastRows && clamp;
  if (useFast) {
    if (a < 0) {
#pragma unroll
      for (int k = 0; k < 10; k++) M[k] = 0.0;
    }
    if (b < 0) {
#pragma unroll
      for (int k = 10; k < 20; k++) M[k] = 0.0;
    }
  }
  double dv1[8], dv2[8], acc1[6], acc2[6];
#pragma unroll
  for (int k = 0; k < 8; k++) { dv1[k] = 0.0; dv2[k] = 0.0; }
  if (a >= 0) { ld4cg(dv + DVS * (size_t)a, dv1); ld4cg(dv + DVS * (size_t)a + 4, dv1 + 4); }
  if (b >= 0) { ld4cg(dv + DVS * (size_t)b, dv2); ld4cg(dv + DVS * (size_t)b + 4, dv2 + 4); }
#pragma unroll
  for (int k = 0; k < 6; k++) { acc1[k] = 0.0; acc2[k] = 0.0; }
  // The chain over the contacts of a pair is sequential; the record of contact c+1 is loaded into a second register
  // set BEFORE the dependent arithmetic of contact c so that its latency hides behind it (the hub variant has no
  // registers to spare and only prefetches it into L2).
  constexpr bool PIPE = !HUB;
  double Q[24], N[24];
  if (PIPE && cnt > 0) {
#pragma unroll
    for (int k = 0; k < 6; k++) ld4(PK0 + 4 * k, Q + 4 * k);
  }
  for (int c = 0; c < cnt; c++) {
    double* PK = S.scP + 24 * (size_t)(start + c);
    if (PIPE) {
      if (c + 1 < cnt) {
#pragma unroll
        for (int k = 0; k < 6; k++) ld4(PK + 24 + 4 * k, N + 4 * k);
        if (c + 2 < cnt) { prefetchL2(PK + 48); prefetchL2(PK + 64); }
      }
    } else {
      if (c + 1 < cnt) { prefetchL2(PK + 24); prefetchL2(PK + 40); }
#pragma unroll
      for (int k = 0; k < 6; k++) ld4(PK + 4 * k, Q + 4 * k);
    }
    d3 dir[3] = {{Q[0], Q[1], Q[2]}, {Q[3], Q[4], Q[5]}, {Q[6], Q[7], Q[8]}};
    d3 r1 = {Q[9], Q[10], Q[11]}, r2 = {Q[12], Q[13], Q[14]};
    double lam[3] = {Q[21], Q[22], Q[23]};
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        d3 jav = vscale(-1, dir[k]), jaw = vcross(dir[k], r1), jbw = vcross(r2, dir[k]);
        if (a >= 0) applyRow(dv1, M[0], M + 1, jav, jaw, lam[k]);
        if (b >= 0) applyRow(dv2, M[10], M + 11, dir[k], jbw, lam[k]);
        if (hubA) applyRow(acc1, M[0], M + 1, jav, jaw, lam[k]);
        if (hubB) applyRow(acc2, M[10], M + 11, dir[k], jbw, lam[k]);
      }
    } else {
      double bb[3] = {Q[15], Q[16], Q[17]};
      double DD[3] = {Q[18], Q[19], Q[20]};
      double den[3], y[3];
#pragma unroll
      for (int k = 0; k < 3; k++) { den[k] = DD[k] + P.compliance; y[k] = 1.0 / den[k]; }  // off the dependent chain
      bool redo = true;
      if (useFast) {
        // branch-free rows (quotFast / clamp*Bits above) on a COPY of the two deltaV; a contact that raised the flag
        // (quotient outside the normal range, bound not a finite non-negative number) is done again below, exactly
        double f1[6], f2[6], g1[6], g2[6], lf[3], dmax = 0;
        unsigned bad = 0;
        unsigned long long mb = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) { f1[k] = dv1[k]; f2[k] = dv2[k]; g1[k] = acc1[k]; g2[k] = acc2[k]; }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          d3 jav = vscale(-1, dir[k]), jaw = vcross(dir[k], r1), jbw = vcross(r2, dir[k]);
          double Jdv = dot6(jav, jaw, f1) + dot6(dir[k], jbw, f2);
          double l = quotFast(DD[k] * lam[k] - P.omega * (bb[k] + Jdv), den[k], y[k], bad);
          l = (k == 0) ? clampNonNegBits(l) : clampAbsBits(l, mu * lf[0], bad);
          lf[k] = l;
          double diff = l - lam[k];
          applyRow(f1, M[0], M + 1, jav, jaw, diff);   // (a pinned side has a zeroed mass packet: its copy stays +0)
          applyRow(f2, M[10], M + 11, dir[k], jbw, diff);
          if (hubA) applyRow(g1, M[0], M + 1, jav, jaw, diff);
          if (hubB) applyRow(g2, M[10], M + 11, dir[k], jbw, diff);
          unsigned long long db = (unsigned long long)__double_as_longlong(diff) & 0x7fffffffffffffffULL;
          mb = db > mb ? db : mb;
        }
        dmax = __longlong_as_double((long long)mb);
        if (!bad) {
          redo = false;
#pragma unroll
          for (int k = 0; k < 6; k++) { dv1[k] = f1[k]; dv2[k] = f2[k]; acc1[k] = g1[k]; acc2[k] = g2[k]; }
#pragma unroll
          for (int k = 0; k < 3; k++) lam[k] = lf[k];
          unsigned long long lb = (unsigned long long)__double_as_longlong(localMax);  // localMax >= 0 or NaN: bit patterns order
          localMax = mb > lb ? dmax : localMax;
        }
      }
      if (redo) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          d3 jav = vscale(-1, dir[k]), jaw = vcross(dir[k], r1), jbw = vcross(r2, dir[k]);
          double Jdv = dot6(jav, jaw, dv1) + dot6(dir[k], jbw, dv2);
          double prev = lam[k];
          double l = divExact(DD[k] * prev - P.omega * (bb[k] + Jdv), den[k], y[k]);
          if (clamp) {
            if (k == 0) l = fmax(0.0, l);
            else {
              double limit = mu * lam[0];
              l = fmax(l, -limit);
              l = fmin(l, limit);
            }
          }
          lam[k] = l;
          double diff = l - prev;
          if (a >= 0) applyRow(dv1, M[0], M + 1, jav, jaw, diff);
          if (b >= 0) applyRow(dv2, M[10], M + 11, dir[k], jbw, diff);
          if (hubA) applyRow(acc1, M[0], M + 1, jav, jaw, diff);
          if (hubB) applyRow(acc2, M[10], M + 11, dir[k], jbw, diff);
          localMax = fmax(localMax, fabs(diff));
        }
      }
      st4(PK + 20, DD[2], lam[0], lam[1], lam[2]);
      if (lastIter) {
        // Contact.updateContactState :385-398
        d3 jaw1 = vcross(dir[1], r1), jbw1 = vcross(r2, dir[1]);
        d3 jaw2 = vcross(dir[2], r1), jbw2 = vcross(r2, dir[2]);
        double w1 = bb[1] + (dot6(vscale(-1, dir[1]), jaw1, dv1) + dot6(dir[1], jbw1, dv2));
        double w2 = bb[2] + (dot6(vscale(-1, dir[2]), jaw2, dv1) + dot6(dir[2], jbw2, dv2));
        int st;
        if (fabs(lam[0]) <= 1e-14) st = AM3D_CS_BROKEN;
        else if (fabs(w1) > P.sliding) st = AM3D_CS_ONEDGE;
        else if (fabs(w2) > P.sliding) st = AM3D_CS_ONEDGE;
        else st = AM3D_CS_CLEAR;
        S.scState[start + c] = st;
      }
    }
    if (PIPE) {
#pragma unroll
      for (int k = 0; k < 24; k++) Q[k] = N[k];
    }
  }
  if (a >= 0) {
    if (!hubA) {
      st4(dv + DVS * (size_t)a, dv1[0], dv1[1], dv1[2], dv1[3]);
      st4(dv + DVS * (size_t)a + 4, dv1[4], dv1[5], 0.0, 0.0);
    } else {
#pragma unroll
      for (int k = 0; k < 6; k++) S.hubDelta[12 * (size_t)p + k] = acc1[k];
    }
  }
  if (b >= 0) {
    if (!hubB) {
      st4(dv + DVS * (size_t)b, dv2[0], dv2[1], dv2[2], dv2[3]);
      st4(dv + DVS * (size_t)b + 4, dv2[4], dv2[5], 0.0, 0.0);
    } else {
#pragma unroll
      for (int k = 0; k < 6; k++) S.hubDelta[12 * (size_t)p + 6 + k] = acc2[k];
    }
  }
  // |delta lambda| >= tolerance somewhere in this group: its scene iterates on (written as "not below" so that a NaN
  // keeps iterating, as Math.max / < do in PGS.java:176,190)
  if (MODE == 1 && P.check) {
    bool moving = !(localMax < P.tolerance);
    unsigned act = __activemask();
    unsigned peers = __match_any_sync(act, scene);           // lanes of this warp that hold a group of the same scene
    bool any = (__ballot_sync(act, moving) & peers) != 0u;
    if (any && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {  // one lane per scene and warp
      int* f = S.sceneState + 2 * P.nScenes + (size_t)scene * MV_SLOTS + (blockIdx.x & (MV_SLOTS - 1));
      if (!__ldcg(f)) *f = 1;
    }
  }
}
// end of an iteration, one thread per scene: count it, and retire the scenes nobody marked as moving
// iterState: [1] every scene is done, [2] largest iteration count, [6] scenes still iterating
__device__ __forceinline__ void sceneIterEnd(int s, const PgsParams& P, int* __restrict__ sceneState, unsigned long long* __restrict__ iterState,
                                             int* partRemaining = nullptr) {
  if (sceneState[s]) return;
  int it = ++sceneState[P.nScenes + s];
  atomicMax(iterState + 2, (unsigned long long)it);
  int* mv = sceneState + 2 * P.nScenes + (size_t)s * MV_SLOTS;
  int moving = 0;
#pragma unroll 4
  for (int k = 0; k < MV_SLOTS; k++) { moving |= __ldcg(mv + k); mv[k] = 0; }
  if (P.check && !moving) {
    sceneState[s] = 1;
    if (atomicAdd(iterState + 6, (unsigned long long)-1LL) == 1ULL) iterState[1] = 1;
    if (partRemaining) atomicSub(partRemaining, 1);
  }
}

// ------------------------------------------------------------------------------------------------
// GIANT groups (two sphere-tree meshes pressed together: 10^3 - 10^4 leaf x leaf contacts in ONE body pair,
// CollisionProcessor.java:975-1056).  Their contacts share both bodies, so Gauss-Seidel prescribes one long chain; what
// can be taken off the chain is everything that does not depend on deltaV.  One WARP per giant group:
//   * prepare: the 32 lanes load 32 consecutive contact records (coalesced) and each lane works out, for its contact,
//     the angular Jacobian halves (d x r1, r2 x d), jinv * (angular half) for both bodies, D + compliance and its
//     reciprocal - the same operations in the same order as the one-thread form, so the results are bit-identical -
//     and parks them in shared memory (60 doubles per contact);
//   * chain: all lanes then walk the 32 contacts in sequence with uniform (broadcast) shared-memory reads: per row
//     only J*deltaV, the lambda update and the two deltaV updates remain (~85 instead of ~200 instructions, no
//     global-memory latency on the chain); lane c keeps contact c's new lambda / state and writes them back coalesced.
// ------------------------------------------------------------------------------------------------
#define GIANT_MIN 65       // contacts from which a group is solved by a warp
#define GIANT_ROW 60       // doubles per prepared contact
#define GIANT_WARPS 4      // warps per CTA of k_pgs_giant
// the chain over one prepared batch (<= 32 contacts in W): every lane runs it (uniform), lane c keeps contact c's results.
// FAST: branch-free row update, returns nonzero if the batch has to be run again with FAST = false.
template <int MODE, bool HUB, bool FAST>
__device__ __forceinline__ unsigned giantChain(const double* __restrict__ W, int n, int lane, bool hasA, bool hasB, bool hubA, bool hubB,
                                               bool clamp, double mu, const double (&M)[20], const PgsParams& P, int lastIter,
                                               double (&dv1)[8], double (&dv2)[8], double (&acc1)[6], double (&acc2)[6],
                                               double& localMax, double (&myLam)[3], double& myD2, int& mySt) {
  unsigned bad = 0;
  unsigned long long maxBits = (unsigned long long)__double_as_longlong(localMax);  // (FAST) |diff| maximum as a bit pattern: NaN stays on top
  for (int c = 0; c < n; c++) {
    const double* R = W + c * GIANT_ROW;
    double lam[3] = {R[57], R[58], R[59]};
    double w12[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      d3 dir = {R[3 * k], R[3 * k + 1], R[3 * k + 2]};
      d3 jav = vscale(-1, dir);
      d3 jaw = {R[9 + 3 * k], R[10 + 3 * k], R[11 + 3 * k]}, jbw = {R[18 + 3 * k], R[19 + 3 * k], R[20 + 3 * k]};
      d3 ta = {R[27 + 3 * k], R[28 + 3 * k], R[29 + 3 * k]}, tb = {R[36 + 3 * k], R[37 + 3 * k], R[38 + 3 * k]};
      double diff;
      if (MODE == 0) diff = lam[k];
      else {
        double Jdv = dot6(jav, jaw, dv1) + dot6(dir, jbw, dv2);
        double prev = lam[k];
        double num = R[48 + k] * prev - P.omega * (R[45 + k] + Jdv);
        double l;
        if (FAST) {
          l = quotFast(num, R[51 + k], R[54 + k], bad);
          l = (k == 0) ? clampNonNegBits(l) : clampAbsBits(l, mu * lam[0], bad);  // (the caller takes FAST only for clamped groups)
        } else {
          l = divExact(num, R[51 + k], R[54 + k]);
          if (clamp) {
            if (k == 0) l = fmax(0.0, l);
            else {
              double limit = mu * lam[0];
              l = fmax(l, -limit);
              l = fmin(l, limit);
            }
          }
        }
        lam[k] = l;
        diff = l - prev;
        if (FAST) {
          unsigned long long db = (unsigned long long)__double_as_longlong(diff) & 0x7fffffffffffffffULL;
          maxBits = db > maxBits ? db : maxBits;
        } else {
          localMax = fmax(localMax, fabs(diff));
        }
      }
      // PGS.updateDeltaVwithLambdai :221-245 with jinv * (angular half) taken from the prepared row.
      // FAST: no test on the body - the prepared rows of a pinned side carry jinv * (angular half) = 0 and its inverse mass
      // is passed as 0, so its deltaV stays +0 exactly as when the update is skipped.
      if (FAST || hasA) {
        double s1 = M[0] * diff;
        dv1[0] = s1 * jav.x + dv1[0]; dv1[1] = s1 * jav.y + dv1[1]; dv1[2] = s1 * jav.z + dv1[2];
        dv1[3] = diff * ta.x + dv1[3]; dv1[4] = diff * ta.y + dv1[4]; dv1[5] = diff * ta.z + dv1[5];
        if (hubA) {
          acc1[0] = s1 * jav.x + acc1[0]; acc1[1] = s1 * jav.y + acc1[1]; acc1[2] = s1 * jav.z + acc1[2];
          acc1[3] = diff * ta.x + acc1[3]; acc1[4] = diff * ta.y + acc1[4]; acc1[5] = diff * ta.z + acc1[5];
        }
      }
      if (FAST || hasB) {
        double s2 = M[10] * diff;
        dv2[0] = s2 * dir.x + dv2[0]; dv2[1] = s2 * dir.y + dv2[1]; dv2[2] = s2 * dir.z + dv2[2];
        dv2[3] = diff * tb.x + dv2[3]; dv2[4] = diff * tb.y + dv2[4]; dv2[5] = diff * tb.z + dv2[5];
        if (hubB) {
          acc2[0] = s2 * dir.x + acc2[0]; acc2[1] = s2 * dir.y + acc2[1]; acc2[2] = s2 * dir.z + acc2[2];
          acc2[3] = diff * tb.x + acc2[3]; acc2[4] = diff * tb.y + acc2[4]; acc2[5] = diff * tb.z + acc2[5];
        }
      }
    }
    if (MODE == 1) {
      int st = 0;
      if (lastIter) {  // Contact.updateContactState :385-398, with the deltaV after this contact's three rows
#pragma unroll
        for (int k = 1; k < 3; k++) {
          d3 dir = {R[3 * k], R[3 * k + 1], R[3 * k + 2]};
          d3 jaw = {R[9 + 3 * k], R[10 + 3 * k], R[11 + 3 * k]}, jbw = {R[18 + 3 * k], R[19 + 3 * k], R[20 + 3 * k]};
          w12[k - 1] = R[45 + k] + (dot6(vscale(-1, dir), jaw, dv1) + dot6(dir, jbw, dv2));
        }
        if (fabs(lam[0]) <= 1e-14) st = AM3D_CS_BROKEN;
        else if (fabs(w12[0]) > P.sliding) st = AM3D_CS_ONEDGE;
        else if (fabs(w12[1]) > P.sliding) st = AM3D_CS_ONEDGE;
        else st = AM3D_CS_CLEAR;
      }
      if (lane == c) { myLam[0] = lam[0]; myLam[1] = lam[1]; myLam[2] = lam[2]; myD2 = R[50]; mySt = st; }
    }
  }
  if (FAST && MODE == 1) localMax = __longlong_as_double((long long)maxBits);
  return bad;
}

// prepare one contact of a giant group: everything of its three rows that does not depend on deltaV, into 60 doubles
__device__ __forceinline__ void giantPrepare(const double* __restrict__ PK, double* __restrict__ R, const double (&M)[20], double compliance) {
  double Q[24];
#pragma unroll
  for (int k = 0; k < 6; k++) ld4(PK + 4 * k, Q + 4 * k);
  d3 r1 = {Q[9], Q[10], Q[11]}, r2 = {Q[12], Q[13], Q[14]};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    d3 dir = {Q[3 * k], Q[3 * k + 1], Q[3 * k + 2]};
    d3 jaw = vcross(dir, r1), jbw = vcross(r2, dir);
    R[3 * k] = dir.x; R[3 * k + 1] = dir.y; R[3 * k + 2] = dir.z;
    R[9 + 3 * k] = jaw.x; R[10 + 3 * k] = jaw.y; R[11 + 3 * k] = jaw.z;
    R[18 + 3 * k] = jbw.x; R[19 + 3 * k] = jbw.y; R[20 + 3 * k] = jbw.z;
    const double* J1 = M + 1;
    const double* J2 = M + 11;
    R[27 + 3 * k] = J1[0] * jaw.x + J1[1] * jaw.y + J1[2] * jaw.z;
    R[28 + 3 * k] = J1[3] * jaw.x + J1[4] * jaw.y + J1[5] * jaw.z;
    R[29 + 3 * k] = J1[6] * jaw.x + J1[7] * jaw.y + J1[8] * jaw.z;
    R[36 + 3 * k] = J2[0] * jbw.x + J2[1] * jbw.y + J2[2] * jbw.z;
    R[37 + 3 * k] = J2[3] * jbw.x + J2[4] * jbw.y + J2[5] * jbw.z;
    R[38 + 3 * k] = J2[6] * jbw.x + J2[7] * jbw.y + J2[8] * jbw.z;
    R[45 + k] = Q[15 + k];                      // b
    R[48 + k] = Q[18 + k];                      // D
    double den = Q[18 + k] + compliance;
    R[51 + k] = den;
    R[54 + k] = 1.0 / den;
    R[57 + k] = Q[21 + k];                      // lambda
  }
}

template <int MODE, bool HUB>
__global__ void __launch_bounds__(32 * GIANT_WARPS)
k_pgs_giant(int gBegin, int gEnd, SolveArrays S, double* __restrict__ dv, PgsParams P, int lastIter,
            unsigned long long* __restrict__ iterState) {
  extern __shared__ __align__(16) double giantSm[];  // [GIANT_WARPS][32][GIANT_ROW]
  if (MODE == 1 && iterState[1]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = gBegin + blockIdx.x * GIANT_WARPS + warp;
  if (p >= gEnd) return;
  int scene = 0;
  if (MODE == 1 && P.check) {
    scene = S.sgScene[p];
    if (S.sceneState[scene]) return;
  }
  double* W = giantSm + (size_t)warp * 32 * GIANT_ROW;
  const int a = S.sgB1[p], b = S.sgB2[p];
  const int start = S.sgStart[p], cnt = S.sgCount[p];
  double M[20];
  const double* Mp = S.sgMass + 20 * (size_t)p;
#pragma unroll
  for (int k = 0; k < 5; k++) ld4(Mp + 4 * k, M + 4 * k);
  const double mu = S.sgMu[p];
  const int fl = S.sgFlags[p];
  const bool clamp = fl & SG_CLAMP;
  const bool hubA = HUB && (fl & SG_HUB1), hubB = HUB && (fl & SG_HUB2);
  // a pinned side (index < 0) is never updated: with its inverse mass and inertia zeroed the unguarded update of the
  // branch-free chain adds +-0 to a deltaV that is +0 (the guarded form reads M only where the side is not pinned)
  if (a < 0) {
#pragma unroll
    for (int k = 0; k < 10; k++) M[k] = 0.0;
  }
  if (b < 0) {
#pragma unroll
    for (int k = 10; k < 20; k++) M[k] = 0.0;
  }
  double dv1[8], dv2[8], acc1[6], acc2[6];
#pragma unroll
  for (int k = 0; k < 8; k++) { dv1[k] = 0.0; dv2[k] = 0.0; }
  if (a >= 0) { ld4cg(dv + DVS * (size_t)a, dv1); ld4cg(dv + DVS * (size_t)a + 4, dv1 + 4); }
  if (b >= 0) { ld4cg(dv + DVS * (size_t)b, dv2); ld4cg(dv + DVS * (size_t)b + 4, dv2 + 4); }
#pragma unroll
  for (int k = 0; k < 6; k++) { acc1[k] = 0.0; acc2[k] = 0.0; }
  double localMax = 0;
  for (int base = 0; base < cnt; base += 32) {
    const int n = min(32, cnt - base);
    double* PK = S.scP + 24 * (size_t)(start + base + lane);
    // ---- prepare: lane i -> contact base + i ----
    if (lane < n) giantPrepare(PK, W + lane * GIANT_ROW, M, P.compliance);
    __syncwarp();
    // ---- chain ----
    double myLam[3] = {0, 0, 0}, myD2 = 0;
    int mySt = 0;
    if (MODE == 1 && clamp && P.fastRows) {  // (unclamped = a contact with an active magnet, PGS.java:119: exact forms)
      double sv1[6], sv2[6], sa1[6], sa2[6], sMax = localMax;  // state at the start of the batch, for the exact re-run
#pragma unroll
      for (int k = 0; k < 6; k++) { sv1[k] = dv1[k]; sv2[k] = dv2[k]; sa1[k] = acc1[k]; sa2[k] = acc2[k]; }
      unsigned bad = giantChain<MODE, HUB, true>(W, n, lane, a >= 0, b >= 0, hubA, hubB, clamp, mu, M, P, lastIter, dv1, dv2, acc1, acc2,
                                                 localMax, myLam, myD2, mySt);
      if (bad) {  // (uniform: every lane ran the same chain)
#pragma unroll
        for (int k = 0; k < 6; k++) { dv1[k] = sv1[k]; dv2[k] = sv2[k]; acc1[k] = sa1[k]; acc2[k] = sa2[k]; }
        localMax = sMax;
        giantChain<MODE, HUB, false>(W, n, lane, a >= 0, b >= 0, hubA, hubB, clamp, mu, M, P, lastIter, dv1, dv2, acc1, acc2, localMax,
                                     myLam, myD2, mySt);
      }
    } else {
      giantChain<MODE, HUB, false>(W, n, lane, a >= 0, b >= 0, hubA, hubB, clamp, mu, M, P, lastIter, dv1, dv2, acc1, acc2, localMax,
                                   myLam, myD2, mySt);
    }
    if (MODE == 1 && lane < n) {
      st4(PK + 20, myD2, myLam[0], myLam[1], myLam[2]);
      if (lastIter) S.scState[start + base + lane] = mySt;
    }
    __syncwarp();
  }
  if (lane == 0) {
    if (a >= 0) {
      if (!hubA) {
        st4(dv + DVS * (size_t)a, dv1[0], dv1[1], dv1[2], dv1[3]);
        st4(dv + DVS * (size_t)a + 4, dv1[4], dv1[5], 0.0, 0.0);
      } else {
#pragma unroll
        for (int k = 0; k < 6; k++) S.hubDelta[12 * (size_t)p + k] = acc1[k];
      }
    }
    if (b >= 0) {
      if (!hubB) {
        st4(dv + DVS * (size_t)b, dv2[0], dv2[1], dv2[2], dv2[3]);
        st4(dv + DVS * (size_t)b + 4, dv2[4], dv2[5], 0.0, 0.0);
      } else {
#pragma unroll
        for (int k = 0; k < 6; k++) S.hubDelta[12 * (size_t)p + 6 + k] = acc2[k];
      }
    }
    if (MODE == 1 && P.check && !(localMax < P.tolerance)) {
      int* f = S.sceneState + 2 * P.nScenes + (size_t)scene * MV_SLOTS + (blockIdx.x & (MV_SLOTS - 1));
      if (!__ldcg(f)) *f = 1;
    }
  }
}
// (Measured and not kept, round 2: a producer warp preparing batch k + 1 into a second shared-memory buffer while a consumer
// warp walks the chain of batch k - no gain: the preparation is ~10 % of a batch; the chain, ~300 instructions per contact of
// which 183 are FP64 at 2 cycles of the pipe each and ~150 cycles per row are a dependent sequence, is what a launch costs.)
// giant groups per phase (they lead their phase: groups are sorted by descending contact count)
__global__ void k_phase_giants(int ng, const int* __restrict__ sgCount, const int* __restrict__ phaseOf, int* __restrict__ phaseGiants) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < ng && sgCount[p] >= GIANT_MIN) atomicAdd(phaseGiants + phaseOf[p], 1);
}

// Fold the per-group deltas of one colour into the hubs' deltaV: one warp per (colour, hub) run, lanes take the
// run's entries strided by 32 (sequential partial sums), then a fixed xor-butterfly: the order never changes.
struct HubRuns {
  const int* runStart;   // [nRuns + 1] into entrySlot
  const int* runBody;    // [nRuns] hub solver body
  const int* entrySlot;  // group * 2 + side, ascending group within a run
};
// sceneDone: the done flags of the scenes during the sweeps (a retired scene's groups wrote no deltas), or nullptr
__device__ __forceinline__ void hubReduceRun(int r, const HubRuns& H, const double* __restrict__ hubDelta, double* __restrict__ dv,
                                             const int* __restrict__ sgScene, const int* __restrict__ sceneDone) {
  int lane = threadIdx.x & 31;
  if (sceneDone && sceneDone[sgScene[H.entrySlot[H.runStart[r]] >> 1]]) return;
  int e0 = H.runStart[r], e1 = H.runStart[r + 1];
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int e = e0 + lane; e < e1; e += 32) {
    int slot = H.entrySlot[e];
    const double* d = hubDelta + 6 * (size_t)slot;
#pragma unroll
    for (int k = 0; k < 6; k++) s[k] = s[k] + __ldcg(d + k);
  }
#pragma unroll
  for (int k = 0; k < 6; k++)
    for (int o = 16; o > 0; o >>= 1) s[k] = s[k] + __shfl_xor_sync(0xffffffffu, s[k], o);
  if (lane == 0) {
    int h = H.runBody[r];
#pragma unroll
    for (int k = 0; k < 6; k++) dv[DVS * (size_t)h + k] = __ldcg(dv + DVS * (size_t)h + k) + s[k];
  }
}
__global__ void k_hub_reduce(int rBegin, int rEnd, HubRuns H, SolveArrays S, double* __restrict__ dv,
                             const unsigned long long* __restrict__ iterState, int mode, int check) {
  if (mode == 1 && iterState[1]) return;
  int r = rBegin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (r < rEnd) hubReduceRun(r, H, S.hubDelta, dv, S.sgScene, (mode == 1 && check) ? S.sceneState : nullptr);
}

#ifndef PGS_MIN_BLOCKS
#define PGS_MIN_BLOCKS 1  // (a register cap for 3 CTAs per SM was measured again in round 2)
#endif
template <int MODE, bool HUB>
__global__ void __launch_bounds__(128, PGS_MIN_BLOCKS)
k_pgs_color(int gBegin, int gEnd, SolveArrays S, double* __restrict__ dv, PgsParams P, int lastIter,
            unsigned long long* __restrict__ iterState) {
  if (MODE == 1 && iterState[1]) return;  // every scene has taken its tolerance exit (PGS.java:190-192)
  int p = gBegin + blockIdx.x * blockDim.x + threadIdx.x;
  if (p < gEnd) pgsGroup<MODE, HUB>(p, S, dv, P, lastIter);
}

// Tail phases of batched scenes.  A body touched by many pairs (the platform of tower25platform.xml: 13) forces as many
// colours, and the last of them hold nothing but ITS pairs: one group per scene.  A launch per such phase is pure latency
// (32 CTAs on 148 SMs, two dependent loads and one short chain each).  They are folded into ONE launch with a thread per
// scene that walks its scene's groups of those phases in phase order - the same Gauss-Seidel sequence, since a scene's
// groups never meet another scene's - and, in the full solve, closes the iteration for its scene (sceneIterEnd) right
// away.  table[(phase - c0) * nScenes + scene] = the scene's group in that phase or -1 (k_tail_table).
__global__ void k_tail_table(int g0, int ng, int c0, int nScenes, const int* __restrict__ sgPhase, const int* __restrict__ sgScene,
                             int* __restrict__ table, int* __restrict__ dupFlag) {
  int p = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ng) return;
  int old = atomicExch(table + (size_t)(sgPhase[p] - c0) * nScenes + sgScene[p], p);
  if (old != -1) *dupFlag = 1;  // two groups of one scene in a tail phase: the caller takes the launch-per-phase path
}
template <int MODE>
__global__ void __launch_bounds__(128)
k_pgs_tail(int nTail, const int* __restrict__ table, SolveArrays S, double* __restrict__ dv, PgsParams P, int lastIter,
           unsigned long long* __restrict__ iterState, int closeIteration) {
  if (MODE == 1 && iterState[1]) return;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.nScenes) return;
  for (int k = 0; k < nTail; k++) {
    int p = table[(size_t)k * P.nScenes + s];
    if (p >= 0) pgsGroup<MODE, false>(p, S, dv, P, lastIter);
  }
  if (MODE == 1 && closeIteration) sceneIterEnd(s, P, S.sceneState, iterState);
}

// The whole solve in ONE cooperative launch: warm-start pass, then `iterations` sweeps, one grid-wide barrier per
// colour (two when the colour has hub runs).  Used when the colours are many and small (batched scenes): thousands
// of tiny launches become grid barriers.  Same Gauss-Seidel sequence as the per-colour launches.
// (Measured without gain on B200, so not kept: an own split arrive/wait barrier without the L1 invalidate, L2 prefetch
// of the next colour's records while waiting, zigzag and ticket-counter distribution of the groups: DESIGN.md.)
template <bool HUB>
__global__ void __launch_bounds__(128)
k_pgs_persistent(int nColors, const int* __restrict__ colorStart, const int* __restrict__ colorRunStart, HubRuns H,
                 SolveArrays S, double* __restrict__ dv, PgsParams P, int iterations, unsigned long long* __restrict__ iterState) {
  cg::grid_group grid = cg::this_grid();
  int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  int warp = tid >> 5, nwarps = stride >> 5;
  for (int c = 0; c < nColors; c++) {
    int g1 = colorStart[c + 1];
    for (int p = colorStart[c] + tid; p < g1; p += stride) pgsGroup<0, HUB>(p, S, dv, P, 0);
    grid.sync();
    if (HUB) {
      int r1 = colorRunStart[c + 1];
      if (r1 > colorRunStart[c]) {
        for (int r = colorRunStart[c] + warp; r < r1; r += nwarps) hubReduceRun(r, H, S.hubDelta, dv, S.sgScene, nullptr);
        grid.sync();
      }
    }
  }
  for (int it = 0; it < iterations; it++) {
    int last = it == iterations - 1;
    for (int c = 0; c < nColors; c++) {
      int g1 = colorStart[c + 1];
      for (int p = colorStart[c] + tid; p < g1; p += stride) pgsGroup<1, HUB>(p, S, dv, P, last);
      grid.sync();
      if (HUB) {
        int r1 = colorRunStart[c + 1];
        if (r1 > colorRunStart[c]) {
          for (int r = colorRunStart[c] + warp; r < r1; r += nwarps) hubReduceRun(r, H, S.hubDelta, dv, S.sgScene, P.check ? S.sceneState : nullptr);
          grid.sync();
        }
      }
    }
    for (int s = tid; s < P.nScenes; s += stride) sceneIterEnd(s, P, S.sceneState, iterState);
    grid.sync();
    if (((volatile unsigned long long*)iterState)[1]) break;
  }
}
// Batched scenes, one THREAD-BLOCK CLUSTER per partition of scenes.  Scenes never interact, so a phase barrier only has to
// cover the scenes that share the barrier: instead of one grid-wide barrier per phase (k_pgs_persistent: every CTA waits
// for the slowest group of ALL scenes, and for the barrier's own latency across 296 CTAs) the scenes are cut into as many
// partitions as clusters fit the GPU, each cluster walks the phases of ITS scenes with the hardware cluster barrier
// (barrier.cluster arrive.release / wait.acquire) and the partitions drift apart freely.  Same Gauss-Seidel sequence per
// scene, so the results are bit-identical to the other sweep forms.
#define PGS_CLUSTER 8
template <bool HUB>
__global__ void __launch_bounds__(128)
k_pgs_cluster(const int* __restrict__ partRange, const int* __restrict__ partSceneStart, int* __restrict__ partRemaining,
              const int* __restrict__ colorStart, const int* __restrict__ colorRunStart, HubRuns H, SolveArrays S,
              double* __restrict__ dv, PgsParams P, int iterations, unsigned long long* __restrict__ iterState) {
  cg::cluster_group cl = cg::this_cluster();
  const int part = blockIdx.x / PGS_CLUSTER;
  const int tid = (int)cl.thread_rank(), stride = (int)cl.num_threads();
  const int warp = tid >> 5, nwarps = stride >> 5;
  const int ph0 = partRange[2 * part], ph1 = partRange[2 * part + 1];
  if (ph1 <= ph0) return;  // (uniform over the cluster)
  for (int c = ph0; c < ph1; c++) {
    int g1 = colorStart[c + 1];
    for (int p = colorStart[c] + tid; p < g1; p += stride) pgsGroup<0, HUB>(p, S, dv, P, 0);
    cl.sync();
    if (HUB) {
      int r1 = colorRunStart[c + 1];
      if (r1 > colorRunStart[c]) {
        for (int r = colorRunStart[c] + warp; r < r1; r += nwarps) hubReduceRun(r, H, S.hubDelta, dv, S.sgScene, nullptr);
        cl.sync();
      }
    }
  }
  const int s0 = partSceneStart[part], s1 = partSceneStart[part + 1];
  for (int it = 0; it < iterations; it++) {
    int last = it == iterations - 1;
    for (int c = ph0; c < ph1; c++) {
      int g1 = colorStart[c + 1];
      for (int p = colorStart[c] + tid; p < g1; p += stride) pgsGroup<1, HUB>(p, S, dv, P, last);
      cl.sync();
      if (HUB) {
        int r1 = colorRunStart[c + 1];
        if (r1 > colorRunStart[c]) {
          for (int r = colorRunStart[c] + warp; r < r1; r += nwarps) hubReduceRun(r, H, S.hubDelta, dv, S.sgScene, P.check ? S.sceneState : nullptr);
          cl.sync();
        }
      }
    }
    for (int s = s0 + tid; s < s1; s += stride) sceneIterEnd(s, P, S.sceneState, iterState, partRemaining + part);
    cl.sync();
    if (P.check && __ldcg(partRemaining + part) <= 0) break;
  }
}
__global__ void k_iter_end(SolveArrays S, PgsParams P, unsigned long long* __restrict__ iterState) {
  if (iterState[1]) return;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < P.nScenes) sceneIterEnd(s, P, S.sceneState, iterState);
}
// contact-iterations of the solve = sum over groups of contacts x iterations of the group's scene -> iterState[4]
__global__ void k_row_updates(int ng, SolveArrays S, int nScenes, unsigned long long* __restrict__ iterState) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = 0;
  if (p < ng) v = (unsigned long long)S.sgCount[p] * (unsigned long long)S.sceneState[nScenes + S.sgScene[p]];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(iterState + 4, v);
}

// hub entries: one per (group, hub side), keyed (colour, hub body, group position) so that a radix sort groups them
// into (colour, hub) runs with ascending group position
__global__ void k_hub_entries(int ng, const int* __restrict__ sgFlags, const int* __restrict__ sgB1, const int* __restrict__ sgB2,
                              const int* __restrict__ phaseOf, const int* __restrict__ scan, int bitsG, int bitsB,
                              unsigned long long* __restrict__ key, int* __restrict__ slot) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ng) return;
  int fl = sgFlags[p];
  int o = scan[p];
  unsigned long long col = (unsigned long long)phaseOf[p];  // key = phase | body | group, bit widths chosen by the host (sum <= 64)
  if (fl & SG_HUB1) { key[o] = (col << (bitsG + bitsB)) | ((unsigned long long)sgB1[p] << bitsG) | (unsigned long long)p; slot[o] = 2 * p; o++; }
  if (fl & SG_HUB2) { key[o] = (col << (bitsG + bitsB)) | ((unsigned long long)sgB2[p] << bitsG) | (unsigned long long)p; slot[o] = 2 * p + 1; }
}
__global__ void k_hub_sides(int ng, const int* __restrict__ sgFlags, int* __restrict__ n) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ng) return;
  int fl = sgFlags[p];
  n[p] = ((fl & SG_HUB1) ? 1 : 0) + ((fl & SG_HUB2) ? 1 : 0);
}
__global__ void k_hub_run_heads(int ne, const unsigned long long* __restrict__ key, int bitsG, int* __restrict__ head) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  head[e] = (e == 0 || (key[e] >> bitsG) != (key[e - 1] >> bitsG)) ? 1 : 0;
}
__global__ void k_hub_run_fill(int ne, const unsigned long long* __restrict__ key, const int* __restrict__ head,
                               const int* __restrict__ scan, int bitsG, int bitsB, int* __restrict__ runStart,
                               int* __restrict__ runBody, int* __restrict__ runColor) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne || !head[e]) return;
  int r = scan[e];
  runStart[r] = e;
  int hb = (int)((key[e] >> bitsG) & ((1ULL << bitsB) - 1));
  runBody[r] = hb;
  runColor[r] = (int)(key[e] >> (bitsG + bitsB));
}

// copy the solution back to the contact sets (set 0 = external, set 1 = internal contacts of collections) and
// count active contacts per external body pair
__global__ void k_post_solve(int nc, SolveArrays S, const int* __restrict__ cbpc0, double* __restrict__ lam0,
                             int* __restrict__ state0, int writeLam0, int* __restrict__ nActive, double* __restrict__ lam1,
                             int* __restrict__ state1) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nc) return;
  int src = S.scSrc[idx];
  int set = src >> 30, i = src & 0x3fffffff;
  const double* PK = S.scP + 24 * (size_t)idx;
  double l0 = PK[21];
  if (set == 0) {
    if (writeLam0) { lam0[3 * i] = l0; lam0[3 * i + 1] = PK[22]; lam0[3 * i + 2] = PK[23]; }
    state0[i] = S.scState[idx];
    if (nActive && cbpc0[i] >= 0 && fabs(l0) > 1e-14) atomicAdd(nActive + cbpc0[i], 1);  // clearBodyPairContacts :213-226
  } else {
    lam1[3 * i] = l0; lam1[3 * i + 1] = PK[22]; lam1[3 * i + 2] = PK[23];
    state1[i] = S.scState[idx];
  }
}

// ------------------------------------------------------------------------------------------------
// integration
// ------------------------------------------------------------------------------------------------
// RigidBody.advanceVelocities :409-417 for awake, unpinned top-level bodies
__global__ void k_advance_velocities(int ns, int nb, const int* __restrict__ alive, const int* __restrict__ parent,
                                     const int* __restrict__ flags, const double* __restrict__ minv,
                                     const double* __restrict__ jinv, const double* __restrict__ force,
                                     const double* __restrict__ torque, const double* __restrict__ dv,
                                     double* __restrict__ v, double* __restrict__ w, double dt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  if (i >= nb ? !alive[i - nb] : parent[i] >= 0) return;
  if (flags[i] & (AM3D_F_PINNED | AM3D_F_SLEEPING)) return;
  d3 vv = vscaleAdd(dt * minv[i], ld3(force + 3 * i), ld3(v + 3 * i));
  vv = vadd(vv, ld3(dv + DVS * (size_t)i));
  d3 dom = mtransform(ldm(jinv + 9 * i), ld3(torque + 3 * i));
  dom = vscale(dt, dom);
  d3 ww = vadd(ld3(w + 3 * i), dom);
  ww = vadd(ww, ld3(dv + DVS * (size_t)i + 3));
  st3(v + 3 * i, vv);
  st3(w + 3 * i, ww);
}

// RigidBody.advancePositions :427-441 (expRodrigues :382-401) + viscous decay (RigidBodySystem.java:428-436)
__global__ void k_advance_positions(int ns, int nb, const int* __restrict__ alive, const int* __restrict__ parent,
                                    const int* __restrict__ flags, double* __restrict__ x, double* __restrict__ R,
                                    const double* __restrict__ v, const double* __restrict__ w, int vstride /* 3, or DVS: move by deltaV */,
                                    const double* __restrict__ jinv0, const double* __restrict__ mA0,
                                    double* __restrict__ jinv, double* __restrict__ mA, double dt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  if (i >= nb ? !alive[i - nb] : parent[i] >= 0) return;
  if (flags[i] & (AM3D_F_PINNED | AM3D_F_SLEEPING)) return;
  d3 vv = ld3(v + (size_t)vstride * i), om = ld3(w + (size_t)vstride * i);
  st3(x + 3 * i, vscaleAdd(dt, vv, ld3(x + 3 * i)));
  double t = vlen(om) * dt;
  m3 Rm = ldm(R + 9 * i);
  if (t > 1e-8) {
    d3 wn = vnormalize(om);
    double c = cos(t), s = sin(t);
    double c1 = 1 - c;
    m3 dR;
    dR.m[0] = c + wn.x * wn.x * c1;
    dR.m[3] = wn.z * s + wn.x * wn.y * c1;
    dR.m[6] = -wn.y * s + wn.x * wn.z * c1;
    dR.m[1] = -wn.z * s + wn.x * wn.y * c1;
    dR.m[4] = c + wn.y * wn.y * c1;
    dR.m[7] = wn.x * s + wn.y * wn.z * c1;
    dR.m[2] = wn.y * s + wn.x * wn.z * c1;
    dR.m[5] = -wn.x * s + wn.y * wn.z * c1;
    dR.m[8] = c + wn.z * wn.z * c1;
    dR = mmul(dR, Rm);
    Rm = mnormalizeCP(dR);
    stm(R + 9 * i, Rm);
  }
  stm(mA + 9 * i, rm0rt(Rm, ldm(mA0 + 9 * i)));
  stm(jinv + 9 * i, rm0rt(Rm, ldm(jinv0 + 9 * i)));
}
__global__ void k_update_inertia(int nb, const int* __restrict__ flags, const double* __restrict__ R,
                                 const double* __restrict__ jinv0, const double* __restrict__ mA0,
                                 double* __restrict__ jinv, double* __restrict__ mA) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  if (flags[i] & AM3D_F_PINNED) return;
  m3 Rm = ldm(R + 9 * i);
  stm(mA + 9 * i, rm0rt(Rm, ldm(mA0 + 9 * i)));
  stm(jinv + 9 * i, rm0rt(Rm, ldm(jinv0 + 9 * i)));
}
__global__ void k_viscous(int ns, int nb, const int* __restrict__ alive, const int* __restrict__ parent,
                          double* __restrict__ v, double* __restrict__ w, double a1, double a2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  if (i >= nb ? !alive[i - nb] : parent[i] >= 0) return;
  st3(v + 3 * i, vscale(a1, ld3(v + 3 * i)));
  st3(w + 3 * i, vscale(a2, ld3(w + 3 * i)));
}

// ------------------------------------------------------------------------------------------------
// accumulateForMerging (BodyPairContact.java:83-121) + external-contact flags for Sleeping.sleep
// ------------------------------------------------------------------------------------------------
__global__ void k_bpc_accumulate(int nbp, const int* __restrict__ bb1, const int* __restrict__ bb2,
                                 const int* __restrict__ bstart, const int* __restrict__ bcount,
                                 const int* __restrict__ nActive, int* __restrict__ alive, const double* __restrict__ lam,
                                 const int* __restrict__ cstate, const int* __restrict__ parent,
                                 const int* __restrict__ flags, const double* __restrict__ x, const double* __restrict__ R,
                                 const double* __restrict__ v, const double* __restrict__ w, const double* __restrict__ bbB,
                                 const int* __restrict__ bbCount, double* __restrict__ mh, int* __restrict__ sh,
                                 int* __restrict__ nm, int* __restrict__ nst, int accum, int* __restrict__ hasExt, double posDt) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbp) return;
  if (nActive[b] == 0) { alive[b] = 0; return; }  // removeEmptyBodyPairContacts :191-208
  int l1 = bb1[b], l2 = bb2[b];
  int a = parent[l1] >= 0 ? parent[l1] : l1, c = parent[l2] >= 0 ? parent[l2] : l2;
  if (accum > 4) accum = 4;
  double m = pairMetric(a, c, flags, x, R, v, w, bbB, bbCount, posDt);
  int n = nm[b];
  if (n < 4) mh[4 * b + n++] = m;
  else { for (int k = 0; k < 3; k++) mh[4 * b + k] = mh[4 * b + k + 1]; mh[4 * b + 3] = m; }
  if (n > accum) { for (int k = 0; k + 1 < n; k++) mh[4 * b + k] = mh[4 * b + k + 1]; n--; }
  nm[b] = n;
  int notOnEdge = 0, act = 0;
  for (int i = bstart[b]; i < bstart[b] + bcount[b]; i++) {
    if (fabs(lam[3 * i]) > 1e-14) {
      act++;
      if (cstate[i] != AM3D_CS_ONEDGE) notOnEdge++;
    }
  }
  int st;
  if (notOnEdge == 0) st = AM3D_CS_ONEDGE;
  else if (notOnEdge >= 2) st = AM3D_CS_CLEAR;
  else if (act == 1) st = AM3D_CS_CLEAR;
  else st = AM3D_CS_ONEDGE;
  n = nst[b];
  if (n < 4) sh[4 * b + n++] = st;
  else { for (int k = 0; k < 3; k++) sh[4 * b + k] = sh[4 * b + k + 1]; sh[4 * b + 3] = st; }
  if (n > accum) { for (int k = 0; k + 1 < n; k++) sh[4 * b + k] = sh[4 * b + k + 1]; n--; }
  nst[b] = n;
}

// stable compaction of the surviving body pairs into next step's "previous" table
__global__ void k_bpc_compact(int nbp, const int* __restrict__ alive, const int* __restrict__ scan,
                              const unsigned long long* __restrict__ key, const int* __restrict__ b1,
                              const int* __restrict__ b2, const double* __restrict__ mh, const int* __restrict__ sh,
                              const int* __restrict__ nm, const int* __restrict__ nst, unsigned long long* __restrict__ okey,
                              int* __restrict__ ob1, int* __restrict__ ob2, double* __restrict__ omh, int* __restrict__ osh,
                              int* __restrict__ onm, int* __restrict__ onst) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbp || !alive[b]) return;
  int o = scan[b];
  okey[o] = key[b]; ob1[o] = b1[b]; ob2[o] = b2[b]; onm[o] = nm[b]; onst[o] = nst[b];
#pragma unroll
  for (int k = 0; k < 4; k++) { omh[4 * o + k] = mh[4 * b + k]; osh[4 * o + k] = sh[4 * b + k]; }
}

// CollisionProcessor.clearBodyPairContacts :213-226 on its own (the step folds it into k_bpc_accumulate)
__global__ void k_bpc_prune(int nbp, const int* __restrict__ nActive, int* __restrict__ alive) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nbp && nActive[b] == 0) alive[b] = 0;
}
__global__ void k_iota(int n, int* __restrict__ a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}
__global__ void k_bpc_gather(int n, const int* __restrict__ src, const unsigned long long* __restrict__ key,
                             const int* __restrict__ b1, const int* __restrict__ b2, const double* __restrict__ mh,
                             const int* __restrict__ sh, const int* __restrict__ nm, const int* __restrict__ nst,
                             unsigned long long* __restrict__ okey, int* __restrict__ ob1, int* __restrict__ ob2,
                             double* __restrict__ omh, int* __restrict__ osh, int* __restrict__ onm, int* __restrict__ onst) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  int b = src[o];
  okey[o] = key[b]; ob1[o] = b1[b]; ob2[o] = b2[b]; onm[o] = nm[b]; onst[o] = nst[b];
#pragma unroll
  for (int k = 0; k < 4; k++) { omh[4 * o + k] = mh[4 * b + k]; osh[4 * o + k] = sh[4 * b + k]; }
}

// Sleeping.sleep :64-72: does a top-level body still have an external, non-pinned body pair?  Evaluated after
// Merging.merge, when pairs that just became internal no longer count.
__global__ void k_has_ext(int nbp, const int* __restrict__ alive, const int* __restrict__ bb1, const int* __restrict__ bb2,
                          const int* __restrict__ parent, const int* __restrict__ flags, int* __restrict__ hasExt) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbp || !alive[b]) return;
  int l1 = bb1[b], l2 = bb2[b];
  if ((flags[l1] & AM3D_F_PINNED) || (flags[l2] & AM3D_F_PINNED)) return;
  int a = parent[l1] >= 0 ? parent[l1] : l1, c = parent[l2] >= 0 ? parent[l2] : l2;
  hasExt[l1] = 1; hasExt[l2] = 1; hasExt[a] = 1; hasExt[c] = 1;
}

__global__ void k_tail_keys(int nt, int base, const unsigned long long* __restrict__ key0, unsigned long long* __restrict__ k,
                            int* __restrict__ v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  k[i] = key0[base + i];
  v[i] = base + i;
}

// bulk velocity pokes from the host (MouseImpulse / Animation style inputs): added to the top-level entity
__global__ void k_add_velocities(int nb, const int* __restrict__ parent, const double* __restrict__ dvl,
                                 const double* __restrict__ dwl, double* __restrict__ v, double* __restrict__ w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  d3 a = ld3(dvl + 3 * i), b = ld3(dwl + 3 * i);
  if (a.x == 0 && a.y == 0 && a.z == 0 && b.x == 0 && b.y == 0 && b.z == 0) return;
  int t = parent[i] >= 0 ? parent[i] : i;
  if (t == i) {
    st3(v + 3 * i, vadd(ld3(v + 3 * i), a));
    st3(w + 3 * i, vadd(ld3(w + 3 * i), b));
  } else {
    atomicAdd(v + 3 * t, a.x); atomicAdd(v + 3 * t + 1, a.y); atomicAdd(v + 3 * t + 2, a.z);
    atomicAdd(w + 3 * t, b.x); atomicAdd(w + 3 * t + 1, b.y); atomicAdd(w + 3 * t + 2, b.z);
  }
}
