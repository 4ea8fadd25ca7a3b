// C ABI (include/am3d.h) and the host-side orchestration of the per-step kernel sequence.
// Sequence = RigidBodySystem.advanceTime (RigidBodySystem.java:102-185); see stepOnce().
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>

#include "am3d_ctx.h"
#include "am3d_detect.cuh"
#include "am3d_step.cuh"

#define LAUNCH(ctx, kernel, grid, block, ...)                  \
  do {                                                         \
    if ((grid) > 0) {                                          \
      kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__); \
      (ctx)->kernelLaunches++;                                 \
    }                                                          \
  } while (0)

template <class F>
static void cubRun(am3d_ctx* c, F f) {
  size_t bytes = 0;
  CK(f(nullptr, bytes));
  c->cubTemp.ensure(bytes + 16);
  CK(f(c->cubTemp.p, bytes));
  c->kernelLaunches++;
}

template <class T>
static void h2d(am3d_ctx* c, DevBuf<T>& d, const T* src, size_t n) {
  d.ensure(n ? n : 1);
  if (n) CK(cudaMemcpyAsync(d.p, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
}
template <class T>
static void h2dv(am3d_ctx* c, DevBuf<T>& d, const std::vector<T>& v) { h2d(c, d, v.data(), v.size()); }

static int readInt(am3d_ctx* c, const int* p) {
  int v;
  CK(cudaMemcpyAsync(&v, p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return v;
}

// exclusive scan of n ints (+ a trailing 0) so that out[n] is the total
static int scanTotal(am3d_ctx* c, DevBuf<int>& in, DevBuf<int>& out, int n) {
  out.ensure(n + 1);
  CK(cudaMemsetAsync(in.p + n, 0, sizeof(int), c->stream));
  cubRun(c, [&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, in.p, out.p, n + 1, c->stream); });
  return readInt(c, out.p + n);
}

static int bitsFor(unsigned long long v) {
  int b = 1;
  while ((v >> b) && b < 63) b++;
  return b;
}

// ------------------------------------------------------------------------------------------------
// scene upload / reset
// ------------------------------------------------------------------------------------------------
static void copyScene(am3d_ctx* c, const am3d_scene* s) {
  auto& H = c->H;
  H.nb = s->n_bodies; H.nsh = s->n_shapes; H.nn = s->n_nodes; H.nsp = s->n_springs; H.nscenes = s->n_scenes > 0 ? s->n_scenes : 1;
  auto cpI = [](std::vector<int>& d, const int32_t* p, size_t n) { d.assign(p, p + n); };
  auto cpD = [](std::vector<double>& d, const double* p, size_t n) { d.assign(p, p + n); };
  size_t nb = H.nb, nsh = H.nsh, nn = H.nn, nsp = H.nsp;
  cpI(H.body_type, s->body_type, nb); cpI(H.body_flags, s->body_flags, nb); cpI(H.body_scene, s->body_scene, nb);
  cpI(H.body_shape_first, s->body_shape_first, nb); cpI(H.body_shape_count, s->body_shape_count, nb);
  cpI(H.body_bb_count, s->body_bb_count, nb);
  cpD(H.body_x, s->body_x, 3 * nb); cpD(H.body_R, s->body_R, 9 * nb); cpD(H.body_v, s->body_v, 3 * nb);
  cpD(H.body_omega, s->body_omega, 3 * nb); cpD(H.body_mass, s->body_mass, nb); cpD(H.body_minv, s->body_minv, nb);
  cpD(H.body_mA0, s->body_mass_angular0, 9 * nb); cpD(H.body_jinv0, s->body_jinv0, 9 * nb);
  cpD(H.body_fric, s->body_friction, nb); cpD(H.body_rest, s->body_restitution, nb); cpD(H.body_bbB, s->body_bbB, 24 * nb);
  cpI(H.shape_type, s->shape_type, nsh); cpI(H.shape_body, s->shape_body, nsh); cpI(H.shape_root, s->shape_tree_root, nsh);
  cpD(H.shape_size, s->shape_size, 3 * nsh); cpD(H.shape_radius, s->shape_radius, nsh); cpD(H.shape_p, s->shape_p, 3 * nsh);
  cpD(H.shape_lR, s->shape_B2C_R, 9 * nsh); cpD(H.shape_lt, s->shape_B2C_t, 3 * nsh);
  cpD(H.node_c, s->node_c, 3 * nn); cpD(H.node_r, s->node_r, nn);
  cpI(H.node_first, s->node_first_child, nn); cpI(H.node_count, s->node_child_count, nn); cpI(H.node_rank, s->node_rank, nn);
  cpI(H.sp_type, s->spring_type, nsp); cpI(H.sp_b1, s->spring_body1, nsp); cpI(H.sp_b2, s->spring_body2, nsp);
  cpD(H.sp_pb1, s->spring_pb1, 3 * nsp); cpD(H.sp_pb2, s->spring_pb2, 3 * nsp); cpD(H.sp_pw, s->spring_pw, 3 * nsp);
  cpD(H.sp_k, s->spring_k, nsp); cpD(H.sp_d, s->spring_d, nsp); cpD(H.sp_l0, s->spring_l0, nsp); cpD(H.sp_ls, s->spring_ls, nsp);
}

static void validateScene(const am3d_scene* s) {
  if (!s || s->n_bodies <= 0 || s->n_shapes <= 0) throw AmError(AM3D_EINVAL, "empty scene");
  if (s->n_bodies >= (1 << 24)) throw AmError(AM3D_EINVAL, "more than 2^24 bodies in one context");
  for (int i = 0; i < s->n_bodies; i++) {
    if (s->body_shape_count[i] < 1 || s->body_shape_count[i] > 255) throw AmError(AM3D_EINVAL, "body shape count out of range (composites are limited to 255 parts)");
    if (s->body_shape_first[i] < 0 || s->body_shape_first[i] + s->body_shape_count[i] > s->n_shapes) throw AmError(AM3D_EINVAL, "body shape range out of bounds");
    if (s->body_scene[i] < 0 || s->body_scene[i] >= (s->n_scenes > 0 ? s->n_scenes : 1)) throw AmError(AM3D_EINVAL, "body scene id out of range");
  }
  for (int i = 0; i < s->n_shapes; i++) {
    if (s->shape_body[i] < 0 || s->shape_body[i] >= s->n_bodies) throw AmError(AM3D_EINVAL, "shape body out of range");
    if (s->shape_type[i] == AM3D_SHAPE_TREE && (s->shape_tree_root[i] < 0 || s->shape_tree_root[i] >= s->n_nodes)) throw AmError(AM3D_EINVAL, "tree root out of range");
  }
  if (s->n_nodes >= (1 << 27)) throw AmError(AM3D_EINVAL, "too many sphere-tree nodes");
}

// (re)initialise every device array from the host copy of the scene: RigidBodySystem.reset() :390-426
static void resetState(am3d_ctx* c) {
  auto& H = c->H;
  int NB = H.nb, NC = NB / 2 + 1, NS = NB + NC;
  c->NB = NB; c->NS = NS; c->NSH = H.nsh; c->NN = H.nn; c->NSP = H.nsp;
  // body arrays padded to NS
  auto padD = [&](const std::vector<double>& v, int w) { std::vector<double> r(v); r.resize((size_t)NS * w, 0.0); return r; };
  auto padI = [&](const std::vector<int>& v, int fill) { std::vector<int> r(v); r.resize(NS, fill); return r; };
  h2dv(c, c->x, padD(H.body_x, 3)); h2dv(c, c->R, padD(H.body_R, 9)); h2dv(c, c->v, padD(H.body_v, 3));
  h2dv(c, c->w, padD(H.body_omega, 3)); h2dv(c, c->mass, padD(H.body_mass, 1)); h2dv(c, c->minv, padD(H.body_minv, 1));
  h2dv(c, c->mA0, padD(H.body_mA0, 9)); h2dv(c, c->jinv0, padD(H.body_jinv0, 9));
  h2dv(c, c->fric, padD(H.body_fric, 1)); h2dv(c, c->rest, padD(H.body_rest, 1)); h2dv(c, c->bbB, padD(H.body_bbB, 24));
  h2dv(c, c->bbCount, padI(H.body_bb_count, 0));
  std::vector<int> fl = H.body_flags;
  for (int i = 0; i < NB; i++) {
    fl[i] &= ~AM3D_F_SLEEPING;
    if (H.body_type[i] == AM3D_BODY_PLANE) fl[i] |= AM3D_F_PINNED;
  }
  h2dv(c, c->flags, padI(fl, 0));
  h2dv(c, c->scene, padI(H.body_scene, 0));
  h2dv(c, c->btype, padI(H.body_type, -1));
  h2dv(c, c->parent, padI(std::vector<int>(NB, -1), -1));
  h2dv(c, c->bShapeFirst, padI(H.body_shape_first, 0)); h2dv(c, c->bShapeCount, padI(H.body_shape_count, 0));
  std::vector<long long> st(NS);
  for (int i = 0; i < NS; i++) st[i] = i;
  h2dv(c, c->stamp, st);
  // world-frame inertia as the loader leaves it (RigidBody.updateRotationalInertiaFromTransformation :311-321)
  std::vector<double> jinv((size_t)NS * 9, 0.0), mA((size_t)NS * 9, 0.0);
  for (int i = 0; i < NB; i++) {
    m3 Rm = ldm(&H.body_R[9 * i]);
    if (!(fl[i] & AM3D_F_PINNED)) {
      stm(&jinv[9 * i], rm0rt(Rm, ldm(&H.body_jinv0[9 * i])));
      stm(&mA[9 * i], rm0rt(Rm, ldm(&H.body_mA0[9 * i])));
    } else {
      stm(&mA[9 * i], ldm(&H.body_mA0[9 * i]));
    }
  }
  h2dv(c, c->jinv, jinv); h2dv(c, c->mA, mA);
  c->force.ensure(3 * NS); c->torque.ensure(3 * NS); c->dv.ensure(6 * NS);
  c->force.zero(3 * NS, c->stream); c->torque.zero(3 * NS, c->stream); c->dv.zero(6 * NS, c->stream);
  c->metricHist.ensure(10 * NS); c->metricHist.zero(10 * NS, c->stream);
  c->metricCount.ensure(NS); c->metricCount.zero(NS, c->stream);
  c->hasExt.ensure(NS); c->hasExt.zero(NS, c->stream);
  c->collAlive.ensure(NC); c->collAlive.zero(NC, c->stream);
  c->bodyBest.ensure(NS); c->bodyMask.ensure(NS);
  // shapes
  h2dv(c, c->shType, H.shape_type); h2dv(c, c->shBody, H.shape_body); h2dv(c, c->shRoot, H.shape_root);
  h2dv(c, c->shSize, H.shape_size); h2dv(c, c->shRadius, H.shape_radius); h2dv(c, c->shP, H.shape_p);
  h2dv(c, c->shLR, H.shape_lR); h2dv(c, c->shLt, H.shape_lt);
  c->shX.ensure(3 * H.nsh); c->shR.ensure(9 * H.nsh); c->shBoundC.ensure(3 * H.nsh); c->shBoundR.ensure(H.nsh);
  h2dv(c, c->ndC, H.node_c); h2dv(c, c->ndR, H.node_r); h2dv(c, c->ndFirst, H.node_first);
  h2dv(c, c->ndCount, H.node_count); h2dv(c, c->ndRank, H.node_rank);
  // broadphase classes: planes / large shapes (tested against everything) / small shapes (grid)
  std::vector<double> rad;
  std::vector<double> br(H.nsh, 0.0);
  for (int s = 0; s < H.nsh; s++) {
    if (H.shape_type[s] == AM3D_SHAPE_BOX) br[s] = H.shape_radius[s];
    else if (H.shape_type[s] == AM3D_SHAPE_TREE) br[s] = H.node_r[H.shape_root[s]];
    if (H.shape_type[s] != AM3D_SHAPE_PLANE) rad.push_back(br[s]);
  }
  double thr = 1e300, maxSmall = 0;
  if (!rad.empty()) {
    std::vector<double> tmp = rad;
    std::nth_element(tmp.begin(), tmp.begin() + tmp.size() / 2, tmp.end());
    thr = 3.0 * tmp[tmp.size() / 2];
  }
  c->hSmall.clear(); c->hLarge.clear(); c->hPlanes.clear();
  std::vector<int> isLarge(H.nsh, 0);
  for (int s = 0; s < H.nsh; s++) {
    if (H.shape_type[s] == AM3D_SHAPE_PLANE) c->hPlanes.push_back(s);
    else if (br[s] > thr) { c->hLarge.push_back(s); isLarge[s] = 1; }
    else { c->hSmall.push_back(s); maxSmall = std::max(maxSmall, br[s]); }
  }
  if (c->hLarge.size() > 4096) {  // degenerate size distribution: fall back to one class
    for (int s : c->hLarge) { c->hSmall.push_back(s); isLarge[s] = 0; maxSmall = std::max(maxSmall, br[s]); }
    c->hLarge.clear();
    std::sort(c->hSmall.begin(), c->hSmall.end());
  }
  c->cellSize = maxSmall > 0 ? 2.0 * maxSmall * 1.0000001 : 1.0;
  c->nSmall = (int)c->hSmall.size(); c->nLarge = (int)c->hLarge.size(); c->nPlanes = (int)c->hPlanes.size();
  h2dv(c, c->smallList, c->hSmall); h2dv(c, c->largeList, c->hLarge); h2dv(c, c->planeList, c->hPlanes);
  h2dv(c, c->shLarge, isLarge);
  c->cellKey.ensure(c->nSmall + 1); c->cellKeySorted.ensure(c->nSmall + 1); c->cellVal.ensure(c->nSmall + 1); c->cellValSorted.ensure(c->nSmall + 1);
  // springs
  h2dv(c, c->spType, H.sp_type); h2dv(c, c->spB1, H.sp_b1); h2dv(c, c->spB2, H.sp_b2);
  h2dv(c, c->spPb1, H.sp_pb1); h2dv(c, c->spPb2, H.sp_pb2); h2dv(c, c->spPw, H.sp_pw);
  h2dv(c, c->spK, H.sp_k); h2dv(c, c->spD, H.sp_d); h2dv(c, c->spL0, H.sp_l0); h2dv(c, c->spLs, H.sp_ls);
  {
    std::vector<std::pair<int, int>> ent;  // (body, spring<<1|side) in spring order
    for (int s = 0; s < H.nsp; s++) {
      ent.push_back({H.sp_b1[s], s << 1});
      if (H.sp_type[s] == AM3D_SPRING_BODYBODY) ent.push_back({H.sp_b2[s], (s << 1) | 1});
    }
    std::stable_sort(ent.begin(), ent.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
    std::vector<int> bodies, start, list;
    for (size_t i = 0; i < ent.size(); i++) {
      if (i == 0 || ent[i].first != ent[i - 1].first) { bodies.push_back(ent[i].first); start.push_back((int)i); }
      list.push_back(ent[i].second);
    }
    start.push_back((int)ent.size());
    c->nSpringBodies = (int)bodies.size();
    h2dv(c, c->spBodies, bodies); h2dv(c, c->spBodyStart, start); h2dv(c, c->spBodyList, list);
  }
  c->counters.ensure(64); c->counters.zero(64, c->stream);
  c->iterState.ensure(8); c->iterState.zero(8, c->stream);
  c->cur.n = 0; c->prev.n = 0; c->bp.n = 0; c->bpPrev.n = 0;
  c->cur.ensure(1024); c->prev.ensure(1024); c->bp.ensure(256); c->bpPrev.ensure(256);
  c->totalSteps = 0;
  c->mergingEvent = false;
  c->nCollections = 0;
  c->nextStamp = NS;
  memset(&c->T, 0, sizeof(c->T));
  CK(cudaStreamSynchronize(c->stream));
}

// ------------------------------------------------------------------------------------------------
// detection
// ------------------------------------------------------------------------------------------------
static void detect(am3d_ctx* c) {
  int nsh = c->NSH;
  std::swap(c->cur, c->prev);  // ContactPool.swapPools (ContactPool.java:60-65): last step's contacts stay readable
  LAUNCH(c, k_shape_update, nblk(nsh), BLK, nsh, c->shType.p, c->shBody.p, c->shRoot.p, c->shRadius.p, c->shLR.p, c->shLt.p,
         c->btype.p, c->x.p, c->R.p, c->ndC.p, c->ndR.p, c->shX.p, c->shR.p, c->shBoundC.p, c->shBoundR.p);
  double inv = 1.0 / c->cellSize;
  if (c->pairKey.cap == 0) {
    size_t cap = (size_t)nsh * 8 + 1024;
    c->pairKey.ensure(cap); c->pairVal.ensure(cap); c->pairKeySorted.ensure(cap); c->pairValSorted.ensure(cap);
  }
  if (c->nSmall > 0) {
    LAUNCH(c, k_cell_keys, nblk(c->nSmall), BLK, c->nSmall, c->smallList.p, c->shBody.p, c->scene.p, c->shBoundC.p, inv,
           c->cellKey.p, c->cellVal.p);
    int endBit = std::min(64, 42 + bitsFor((unsigned long long)c->H.nscenes));
    cubRun(c, [&](void* t, size_t& b) {
      return cub::DeviceRadixSort::SortPairs(t, b, c->cellKey.p, c->cellKeySorted.p, c->cellVal.p, c->cellValSorted.p, c->nSmall, 0, endBit, c->stream);
    });
  }
  int np = 0;
  for (int attempt = 0; attempt < 3; attempt++) {
    CK(cudaMemsetAsync(c->counters.p, 0, sizeof(int), c->stream));
    PairCtx PC{c->shBody.p, c->bShapeFirst.p, c->parent.p, c->flags.p, c->scene.p, c->stamp.p, c->shBoundC.p, c->shBoundR.p,
               c->pairKey.p, c->pairVal.p, c->counters.p, (int)std::min<size_t>(c->pairKey.cap, 0x7fffffff)};
    if (c->nSmall > 0)
      LAUNCH(c, k_pairs_grid, nblk(c->nSmall), BLK, c->nSmall, c->cellKeySorted.p, c->cellValSorted.p, inv, PC);
    if (c->nLarge > 0 || c->nPlanes > 0)
      LAUNCH(c, k_pairs_special, nblk(nsh), BLK, nsh, c->shType.p, c->shLarge.p, c->nLarge, c->largeList.p, c->nPlanes,
             c->planeList.p, c->shSize.p, c->shRadius.p, PC);
    np = readInt(c, c->counters.p);
    if ((size_t)np <= c->pairKey.cap) break;
    size_t cap = (size_t)np + np / 4 + 1024;
    c->pairKey.ensure(cap); c->pairVal.ensure(cap); c->pairKeySorted.ensure(cap); c->pairValSorted.ensure(cap);
  }
  c->nPairs = np;
  c->T.n_pairs = np;
  int nc = 0;
  if (np > 0) {
    int endBit = std::min(64, 40 + bitsFor((unsigned long long)c->NB));
    cubRun(c, [&](void* t, size_t& b) {
      return cub::DeviceRadixSort::SortPairs(t, b, c->pairKey.p, c->pairKeySorted.p, c->pairVal.p, c->pairValSorted.p, np, 0, endBit, c->stream);
    });
    c->pairType.ensure(np + 1); c->pairCap.ensure(np + 1); c->pairSlot.ensure(np + 1); c->pairCount.ensure(np + 1); c->pairOut.ensure(np + 1);
    LAUNCH(c, k_pair_classify, nblk(np), BLK, np, c->pairValSorted.p, c->shType.p, c->pairType.p, c->pairCap.p);
    TreeCtx TC{c->shType.p, c->shRoot.p, c->shSize.p, c->shRadius.p, c->shP.p, c->shX.p, c->shR.p, c->ndC.p, c->ndR.p, c->ndFirst.p, c->ndCount.p};
    HitOut HO{c->hitPos.p, c->hitNrm.p, c->hitViol.p, c->hitMeta.p};
    bool haveTrees = c->NN > 0;
    CK(cudaMemsetAsync(c->counters.p + 1, 0, sizeof(int), c->stream));
    if (haveTrees)
      LAUNCH(c, k_narrow_tree<false>, nblk(np, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, np, c->pairValSorted.p, c->pairType.p,
             c->pairSlot.p, TC, HO, c->pairCap.p, c->counters.p + 1);
    int nslots = scanTotal(c, c->pairCap, c->pairSlot, np);
    c->nSlots = nslots;
    c->hitPos.ensure(3 * (size_t)nslots + 3); c->hitNrm.ensure(3 * (size_t)nslots + 3); c->hitViol.ensure((size_t)nslots + 1);
    c->hitMeta.ensure(4 * (size_t)nslots + 4);
    HO = HitOut{c->hitPos.p, c->hitNrm.p, c->hitViol.p, c->hitMeta.p};
    CK(cudaMemsetAsync(c->pairCount.p, 0, (np + 1) * sizeof(int), c->stream));
    LAUNCH(c, k_narrow_box, nblk(np, 128), 128, np, c->pairValSorted.p, c->pairType.p, c->pairSlot.p, c->shSize.p, c->shRadius.p,
           c->shX.p, c->shR.p, HO, c->pairCount.p);
    if (haveTrees)
      LAUNCH(c, k_narrow_tree<true>, nblk(np, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, np, c->pairValSorted.p, c->pairType.p,
             c->pairSlot.p, TC, HO, c->pairCount.p, c->counters.p + 1);
    nc = scanTotal(c, c->pairCount, c->pairOut, np);
    if (haveTrees && readInt(c, c->counters.p + 1)) throw AmError(AM3D_ECAPACITY, "sphere-tree traversal stack overflow");
    c->cur.ensure(nc + 1);
    ContactOut CO{c->cur.b1.p, c->cur.b2.p, c->cur.s1.p, c->cur.s2.p, c->cur.bv1.p, c->cur.bv2.p, c->cur.info.p, c->cur.leaf.p,
                  c->cur.state.p, c->cur.isNew.p, c->cur.key0.p, c->cur.key1.p, c->cur.pW.p, c->cur.nW.p, c->cur.t1W.p,
                  c->cur.t2W.p, c->cur.pB1.p, c->cur.nB1.p, c->cur.t1B1.p, c->cur.t2B1.p, c->cur.viol.p, c->cur.prevViol.p,
                  c->cur.lam.p, c->cur.lamWarm.p};
    LAUNCH(c, k_contact_set, nblk(np, 128), 128, np, c->pairKeySorted.p, c->pairValSorted.p, c->pairSlot.p, c->pairCount.p,
           c->pairOut.p, c->shBody.p, c->x.p, c->R.p, c->hitPos.p, c->hitNrm.p, c->hitViol.p, c->hitMeta.p, CO);
  }
  c->cur.n = nc;
  c->T.n_contacts = nc;
}

// updateBodyPairContacts (CollisionProcessor.java:145-164): body pairs of this step, histories carried over
static void buildBodyPairs(am3d_ctx* c) {
  int nc = c->cur.n;
  int nbp = 0;
  if (nc > 0) {
    c->tmpI0.ensure(nc + 1); c->tmpI1.ensure(nc + 1);
    LAUNCH(c, k_bpc_heads, nblk(nc), BLK, nc, c->cur.key0.p, c->cur.b1.p, c->cur.b2.p, c->flags.p, c->tmpI0.p);
    nbp = scanTotal(c, c->tmpI0, c->tmpI1, nc);
    c->bp.ensure(nbp + 1);
    LAUNCH(c, k_bpc_fill, nblk(nc), BLK, nc, c->cur.key0.p, c->tmpI0.p, c->tmpI1.p, c->cur.b1.p, c->cur.b2.p, c->flags.p,
           c->cur.bpc.p, c->bp.key.p, c->bp.start.p, c->bp.b1.p, c->bp.b2.p);
    LAUNCH(c, k_bpc_match, nblk(nbp), BLK, nbp, nc, c->bp.key.p, c->bp.start.p, c->bp.count.p, c->bp.b1.p, c->bp.b2.p,
           c->bp.nActive.p, c->bp.metricHist.p, c->bp.stateHist.p, c->bp.nMetric.p, c->bp.nState.p, c->bp.alive.p,
           c->bpPrev.n, c->bpPrev.key.p, c->bpPrev.b1.p, c->bpPrev.b2.p, c->bpPrev.metricHist.p, c->bpPrev.stateHist.p,
           c->bpPrev.nMetric.p, c->bpPrev.nState.p);
  }
  c->bp.n = nbp;
}

static void warmStart(am3d_ctx* c) {
  int nbp = c->bp.n;
  if (nbp == 0) return;
  WarmCtx W{c->cur.b1.p, c->cur.b2.p, c->cur.s1.p, c->cur.s2.p, c->cur.leaf.p, c->cur.key0.p, c->cur.key1.p, c->cur.pB1.p,
            c->cur.lam.p, c->cur.lamWarm.p, c->cur.prevViol.p, c->cur.isNew.p,
            c->prev.n, c->prev.key0.p, c->prev.key1.p, c->prev.b1.p, c->prev.leaf.p, c->prev.pB1.p, c->prev.viol.p, c->prev.lam.p,
            c->btype.p, c->shType.p, c->x.p, c->R.p, c->ndRank.p};
  LAUNCH(c, k_warm_start, nblk(nbp, 128), 128, nbp, c->bp.start.p, c->bp.count.p, c->bp.b1.p, c->bp.b2.p, W);
}

// ------------------------------------------------------------------------------------------------
// solve
// ------------------------------------------------------------------------------------------------
static SolveArrays solveArrays(am3d_ctx* c) {
  return SolveArrays{c->sgB1.p, c->sgB2.p, c->sgStart.p, c->sgCount.p, c->sgFlags.p, c->sgBpc.p, c->sgMass.p, c->sgMu.p,
                     c->scD.p, c->scR.p, c->scB.p, c->scDiag.p, c->scLam.p, c->scSrc.p, c->scState.p};
}

// colour the groups (body pairs) so that no two groups of a colour share a non-pinned solver body
static void colourGroups(am3d_ctx* c, int ng, int inCollection) {
  c->grpColor.ensure(ng + 1); c->grpPrio.ensure(ng + 1); c->grpSb1.ensure(ng + 1); c->grpSb2.ensure(ng + 1);
  c->grpPos.ensure(ng + 1); c->grpOrder.ensure(ng + 1); c->grpKey.ensure(ng + 1); c->grpKeySorted.ensure(ng + 1); c->grpVal.ensure(ng + 1);
  LAUNCH(c, k_grp_init, nblk(ng), BLK, ng, c->bp.b1.p, c->bp.b2.p, c->parent.p, c->flags.p, inCollection, c->grpSb1.p,
         c->grpSb2.p, c->grpPrio.p, c->grpColor.p);
  CK(cudaMemsetAsync(c->bodyBest.p, 0, c->NS * sizeof(unsigned long long), c->stream));
  CK(cudaMemsetAsync(c->bodyMask.p, 0, c->NS * sizeof(unsigned long long), c->stream));
  int page = 0;
  int maxPages = 64;
  while (true) {
    int remaining = 1, deferred = 0;
    int rounds = 0;
    while (remaining > 0) {
      // a few rounds per host read-back
      for (int r = 0; r < 4; r++) {
        CK(cudaMemsetAsync(c->counters.p + 2, 0, sizeof(int), c->stream));
        LAUNCH(c, k_color_bid, nblk(ng), BLK, ng, c->grpSb1.p, c->grpSb2.p, c->grpPrio.p, c->grpColor.p, c->bodyBest.p);
        LAUNCH(c, k_color_assign, nblk(ng), BLK, ng, page, c->grpSb1.p, c->grpSb2.p, c->grpPrio.p, c->grpColor.p, c->bodyBest.p,
               c->bodyMask.p, c->counters.p + 2, c->counters.p + 3);
      }
      remaining = readInt(c, c->counters.p + 2);
      if (++rounds > 100000) throw AmError(AM3D_ECUDA, "colouring did not converge");
    }
    deferred = readInt(c, c->counters.p + 3);
    if (deferred == 0) break;
    if (++page >= maxPages) throw AmError(AM3D_ECAPACITY, "more than 4096 colours needed");
    CK(cudaMemsetAsync(c->counters.p + 3, 0, sizeof(int), c->stream));
    CK(cudaMemsetAsync(c->bodyMask.p, 0, c->NS * sizeof(unsigned long long), c->stream));
    LAUNCH(c, k_color_next_page, nblk(ng), BLK, ng, page - 1, c->grpColor.p);
  }
  int maxColors = (page + 1) * 64;
  c->colorHist.ensure(maxColors + 1);
  CK(cudaMemsetAsync(c->colorHist.p, 0, (maxColors + 1) * sizeof(int), c->stream));
  LAUNCH(c, k_color_sortkey, nblk(ng), BLK, ng, c->grpColor.p, c->grpKey.p, c->grpVal.p, c->colorHist.p);
  int endBit = 32 + bitsFor((unsigned long long)maxColors);
  cubRun(c, [&](void* t, size_t& b) {
    return cub::DeviceRadixSort::SortPairs(t, b, c->grpKey.p, c->grpKeySorted.p, c->grpVal.p, c->grpOrder.p, ng, 0, endBit, c->stream);
  });
  std::vector<int> hist(maxColors);
  CK(cudaMemcpyAsync(hist.data(), c->colorHist.p, maxColors * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->colorStart.clear();
  int acc = 0;
  for (int k = 0; k < maxColors; k++) {
    if (hist[k] == 0) continue;  // colours are dense within a page but a later page may leave gaps
    c->colorStart.push_back(acc);
    acc += hist[k];
  }
  c->colorStart.push_back(acc);
  c->nColors = (int)c->colorStart.size() - 1;
  c->nGroups = ng;
}

// PGS.solve (PGS.java:73-194) over the current external contacts (full solve, parents as solver bodies)
static void solveFull(am3d_ctx* c, double dt, bool timeIt) {
  int nc = c->cur.n, ng = c->bp.n;
  c->T.pgs_iterations = 0;
  c->T.pgs_colors = 0;
  c->T.pgs_kernel_time = 0;
  if (nc == 0 || ng == 0) return;
  const am3d_params& P = c->P;
  colourGroups(c, ng, 0);
  c->sgB1.ensure(ng + 1); c->sgB2.ensure(ng + 1); c->sgStart.ensure(ng + 2); c->sgCount.ensure(ng + 2); c->sgFlags.ensure(ng + 1);
  c->sgBpc.ensure(ng + 1); c->sgMass.ensure(20 * (size_t)ng + 20); c->sgMu.ensure(ng + 1);
  c->scD.ensure(9 * (size_t)nc + 9); c->scR.ensure(6 * (size_t)nc + 6); c->scB.ensure(3 * (size_t)nc + 3);
  c->scDiag.ensure(3 * (size_t)nc + 3); c->scLam.ensure(3 * (size_t)nc + 3); c->scSrc.ensure(nc + 1); c->scState.ensure(nc + 1);
  SolveArrays S = solveArrays(c);
  LAUNCH(c, k_group_setup, nblk(ng), BLK, ng, c->grpOrder.p, c->grpSb1.p, c->grpSb2.p, c->bp.b1.p, c->bp.b2.p, c->bp.count.p,
         c->minv.p, c->jinv.p, c->fric.p, c->flags.p, P.friction_override, P.friction, S, c->grpPos.p);
  scanTotal(c, c->sgCount, c->sgStart, ng);
  double feedback = P.feedback_stiffness;
  LAUNCH(c, k_assemble, nblk(nc, 128), 128, nc, c->cur.bpc.p, c->bp.start.p, c->grpPos.p, c->sgStart.p, c->cur.b1.p, c->cur.b2.p,
         c->parent.p, 0, c->cur.pW.p, c->cur.nW.p, c->cur.t1W.p, c->cur.t2W.p, c->cur.pB1.p, c->cur.nB1.p, c->cur.t1B1.p,
         c->cur.t2B1.p, c->cur.viol.p, c->cur.lam.p, c->cur.state.p, c->x.p, c->R.p, c->v.p, c->w.p, c->force.p, c->torque.p,
         c->minv.p, c->jinv.p, c->rest.p, dt, feedback, P.restitution_override, P.restitution, S);
  CK(cudaMemsetAsync(c->iterState.p, 0, 8 * sizeof(unsigned long long), c->stream));
  PgsParams PP{P.omega, P.enable_compliance ? P.compliance : 0.0, P.tolerance, P.sliding_threshold};
  if (timeIt) CK(cudaEventRecord(c->ev[10], c->stream));
  for (int k = 0; k < c->nColors; k++) {
    int g0 = c->colorStart[k], g1 = c->colorStart[k + 1];
    LAUNCH(c, k_pgs_color<0>, nblk(g1 - g0, 128), 128, g0, g1, S, c->dv.p, PP, 0, c->iterState.p);
  }
  for (int it = 0; it < P.iterations; it++) {
    int last = it == P.iterations - 1;
    for (int k = 0; k < c->nColors; k++) {
      int g0 = c->colorStart[k], g1 = c->colorStart[k + 1];
      LAUNCH(c, k_pgs_color<1>, nblk(g1 - g0, 128), 128, g0, g1, S, c->dv.p, PP, last, c->iterState.p);
      c->solveLaunches++;
    }
    LAUNCH(c, k_iter_end, 1, 1, c->iterState.p, P.tolerance, 1);
  }
  if (timeIt) CK(cudaEventRecord(c->ev[11], c->stream));
  CK(cudaMemsetAsync(c->bp.nActive.p, 0, ng * sizeof(int), c->stream));
  LAUNCH(c, k_post_solve, nblk(nc), BLK, nc, S, c->cur.bpc.p, c->cur.lam.p, c->cur.state.p, c->bp.nActive.p);
  unsigned long long st[4];
  CK(cudaMemcpyAsync(st, c->iterState.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->T.pgs_iterations = (int)st[2];
  c->T.pgs_colors = c->nColors;
  if (timeIt) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev[10], c->ev[11]));
    c->T.pgs_kernel_time = ms * 1e-3;
  }
  c->rowUpdates += 3.0 * nc * (double)st[2];
  c->solveSeconds += c->T.pgs_kernel_time;
}

// ------------------------------------------------------------------------------------------------
// the step
// ------------------------------------------------------------------------------------------------
static void applyExternalForces(am3d_ctx* c) {
  const am3d_params& P = c->P;
  double theta = P.gravity_angle_deg / 180.0 * M_PI;
  double gx = P.gravity_amount * cos(theta), gy = P.gravity_amount * sin(theta);
  LAUNCH(c, k_clear_gravity, nblk(c->NS), BLK, c->NS, c->NB, c->collAlive.p, c->parent.p, c->mass.p, c->x.p, c->v.p, c->w.p,
         c->force.p, c->torque.p, c->dv.p, P.use_gravity, gx, gy);
  if (P.springs_enabled && c->nSpringBodies > 0)
    LAUNCH(c, k_springs, nblk(c->nSpringBodies, 64), 64, c->nSpringBodies, c->spBodies.p, c->spBodyStart.p, c->spBodyList.p,
           c->spType.p, c->spB1.p, c->spB2.p, c->spPb1.p, c->spPb2.p, c->spPw.p, c->spK.p, c->spD.p, c->spL0.p, c->spLs.p,
           P.spring_k_mod, P.spring_d_mod, c->parent.p, c->x.p, c->R.p, c->v.p, c->w.p, c->force.p, c->torque.p);
}

static float evMs(am3d_ctx* c, int a, int b) {
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]);
  return ms;
}

static void stepOnce(am3d_ctx* c, double dt) {
  const am3d_params& P = c->P;
  c->totalSteps++;
  CK(cudaEventRecord(c->ev[0], c->stream));
  applyExternalForces(c);                       // clearBodies + applyExternalForces  (:108-116)
  CK(cudaEventRecord(c->ev[1], c->stream));
  detect(c);                                    // updateContactsMap + collisionDetection (:119-120)
  buildBodyPairs(c);                            // updateBodyPairContacts (:121)
  CK(cudaEventRecord(c->ev[2], c->stream));
  warmStart(c);                                 // warmStart(false) (:124)
  CK(cudaEventRecord(c->ev[3], c->stream));
  if (P.enable_sleeping) {                      // sleeping.wake() (:128)
    LAUNCH(c, k_wake_pairs, nblk(c->bp.n), BLK, c->bp.n, c->bp.b1.p, c->bp.b2.p, c->parent.p, c->flags.p, c->metricCount.p);
    LAUNCH(c, k_wake_springs, nblk(c->NSP), BLK, c->NSP, c->spType.p, c->spB1.p, c->spB2.p, c->parent.p, c->flags.p, c->metricCount.p);
  }
  // updateInCollections / unmerge (:131-143): no collections can exist while merging is off
  if (P.enable_merging) throw AmError(AM3D_EUNSUPPORTED, "merging is not built in this version of the library yet");
  CK(cudaEventRecord(c->ev[4], c->stream));
  if (!P.warm_start && c->cur.n)                // noWarmStart (:464-470); redoWarmStart is the identity without a single sweep
    CK(cudaMemsetAsync(c->cur.lam.p, 0, 3 * (size_t)c->cur.n * sizeof(double), c->stream));
  solveFull(c, dt, true);                       // solveLCP (:151)
  CK(cudaEventRecord(c->ev[5], c->stream));
  // clearBodyPairContacts (:152) is folded into k_post_solve (nActive) + k_bpc_accumulate
  LAUNCH(c, k_advance_velocities, nblk(c->NS), BLK, c->NS, c->NB, c->collAlive.p, c->parent.p, c->flags.p, c->minv.p, c->jinv.p,
         c->force.p, c->torque.p, c->dv.p, c->v.p, c->w.p, dt);                                  // (:155)
  CK(cudaMemsetAsync(c->hasExt.p, 0, c->NS * sizeof(int), c->stream));
  int nbp = c->bp.n;
  if (nbp > 0) {
    if (c->cur.n == 0) CK(cudaMemsetAsync(c->bp.nActive.p, 0, nbp * sizeof(int), c->stream));
    LAUNCH(c, k_bpc_accumulate, nblk(nbp, 128), 128, nbp, c->bp.b1.p, c->bp.b2.p, c->bp.start.p, c->bp.count.p, c->bp.nActive.p,
           c->bp.alive.p, c->cur.lam.p, c->cur.state.p, c->parent.p, c->flags.p, c->x.p, c->R.p, c->v.p, c->w.p, c->bbB.p,
           c->bbCount.p, c->bp.metricHist.p, c->bp.stateHist.p, c->bp.nMetric.p, c->bp.nState.p, P.step_accum_merging,
           c->hasExt.p);                                                                         // (:158)
  }
  LAUNCH(c, k_advance_positions, nblk(c->NS), BLK, c->NS, c->NB, c->collAlive.p, c->parent.p, c->flags.p, c->x.p, c->R.p, c->v.p,
         c->w.p, c->jinv0.p, c->mA0.p, c->jinv.p, c->mA.p, dt);                                   // (:160)
  // surviving body pairs become next step's lookup table
  int nAlive = 0;
  if (nbp > 0) {
    c->tmpI1.ensure(nbp + 2);
    nAlive = scanTotal(c, c->bp.alive, c->tmpI1, nbp);
    c->bpPrev.ensure(nAlive + 1);
    LAUNCH(c, k_bpc_compact, nblk(nbp), BLK, nbp, c->bp.alive.p, c->tmpI1.p, c->bp.key.p, c->bp.b1.p, c->bp.b2.p, c->bp.metricHist.p,
           c->bp.stateHist.p, c->bp.nMetric.p, c->bp.nState.p, c->bpPrev.key.p, c->bpPrev.b1.p, c->bpPrev.b2.p,
           c->bpPrev.metricHist.p, c->bpPrev.stateHist.p, c->bpPrev.nMetric.p, c->bpPrev.nState.p);
  }
  c->bpPrev.n = nAlive;
  CK(cudaEventRecord(c->ev[6], c->stream));
  if (P.enable_sleeping)                                                                         // (:170)
    LAUNCH(c, k_sleep, nblk(c->NS), BLK, c->NS, c->NB, c->collAlive.p, c->parent.p, c->flags.p, c->hasExt.p, c->metricHist.p,
           c->metricCount.p, c->x.p, c->R.p, c->v.p, c->w.p, c->bbB.p, c->bbCount.p, P.sleep_step_accum, P.sleep_threshold);
  LAUNCH(c, k_viscous, nblk(c->NS), BLK, c->NS, c->NB, c->collAlive.p, c->parent.p, c->v.p, c->w.p, P.viscous_linear, P.viscous_angular);  // (:173)
  CK(cudaEventRecord(c->ev[7], c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->T.detection = evMs(c, 1, 2) * 1e-3;
  c->T.warmstart = evMs(c, 2, 3) * 1e-3;
  c->T.lcp_solve = evMs(c, 4, 5) * 1e-3;
  c->T.update_collections = c->T.contact_ordering = c->T.single_it_pgs = 0;
  c->T.merging = evMs(c, 5, 6) * 1e-3;
  c->T.merging_build = c->T.unmerging = c->T.unmerging_build = 0;
  c->T.compute_time = evMs(c, 0, 7) * 1e-3;
  c->T.n_bodies = c->NB - 0;
  c->T.n_collections = c->nCollections;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
#define API_BEGIN(ctx)                      \
  if (!(ctx)) return AM3D_EINVAL;           \
  try {                                     \
    cudaSetDevice((ctx)->device);
#define API_END(ctx)                        \
    return AM3D_OK;                         \
  } catch (const AmError& e) {              \
    (ctx)->lastError = e.msg;               \
    return e.code;                          \
  } catch (const std::exception& e) {       \
    (ctx)->lastError = e.what();            \
    return AM3D_ECUDA;                      \
  }

static void checkParams(const am3d_params* p) {
  if (p->shuffle) throw AmError(AM3D_EUNSUPPORTED, "shuffle is not supported");
  if (p->enable_post_stabilization) throw AmError(AM3D_EUNSUPPORTED, "post-stabilisation is not supported");
  if (p->collection_cd != 0) throw AmError(AM3D_EUNSUPPORTED, "only the brute-force collection collision mode is supported");
  if (p->use_coriolis) throw AmError(AM3D_EUNSUPPORTED, "Coriolis term is not supported");
  if (p->merge_cycle_condition) throw AmError(AM3D_EUNSUPPORTED, "cycle merge condition is not supported");
  if (p->metric_position_level) throw AmError(AM3D_EUNSUPPORTED, "position-level metric is not supported");
  if (p->iterations < 1 || p->iterations_in_collection < 1) throw AmError(AM3D_EINVAL, "iterations must be >= 1");
  if (p->step_accum_merging > 4 || p->step_accum_unmerging > 4 || p->step_accum_merging < 0) throw AmError(AM3D_EUNSUPPORTED, "accumulation windows above 4 steps are not supported");
  if (p->sleep_step_accum > 10 || p->sleep_step_accum < 0) throw AmError(AM3D_EUNSUPPORTED, "sleep accumulation above 10 steps is not supported");
  if (p->steps_between_merge < 1) throw AmError(AM3D_EINVAL, "steps_between_merge must be >= 1");
}

extern "C" {

const char* am3d_version(void) { return "am3d 0.1 (sm_100a)"; }

void am3d_default_params(am3d_params* p) {
  memset(p, 0, sizeof(*p));
  p->warm_start = 1; p->enable_compliance = 1; p->iterations = 30; p->iterations_in_collection = 1;
  p->feedback_stiffness = 0.5; p->compliance = 1e-3; p->restitution = 0.5; p->friction = 0.1; p->tolerance = 1e-5;
  p->omega = 1.0; p->sliding_threshold = 0.01;
  p->use_gravity = 1; p->springs_enabled = 1; p->gravity_amount = 1.0; p->gravity_angle_deg = 90.0;
  p->viscous_linear = 1.0; p->viscous_angular = 1.0; p->spring_k_mod = 1.0; p->spring_d_mod = 1.0;
  p->enable_merging = 1; p->merge_pinned = 1; p->merge_stable_contact = 1; p->merge_let_it_breathe = 1;
  p->enable_unmerging = 1; p->unmerge_friction = 1; p->unmerge_normal = 1; p->unmerge_relative_motion = 1;
  p->update_contacts_in_collections = 1; p->organize_contacts = 1;
  p->step_accum_merging = 3; p->step_accum_unmerging = 3; p->steps_between_merge = 10;
  p->threshold_merge = 1e-2; p->threshold_unmerge = 2e-2; p->threshold_breath = 1e-5;
  p->enable_sleeping = 1; p->sleep_step_accum = 10; p->sleep_threshold = 1e-5;
}

int am3d_create(int device, am3d_ctx** out) {
  if (!out) return AM3D_EINVAL;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return AM3D_ENOGPU;  // no CPU fallback
  am3d_ctx* c = new am3d_ctx();
  c->device = device;
  try {
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 16; i++) CK(cudaEventCreate(&c->ev[i]));
    c->evCreated = true;
    am3d_default_params(&c->P);
  } catch (const AmError& e) {
    delete c;
    return e.code;
  }
  *out = c;
  return AM3D_OK;
}

int am3d_destroy(am3d_ctx* c) {
  if (!c) return AM3D_EINVAL;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->evCreated) for (int i = 0; i < 16; i++) cudaEventDestroy(c->ev[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return AM3D_OK;
}

const char* am3d_last_error(const am3d_ctx* c) { return c ? c->lastError.c_str() : "null context"; }

int am3d_upload_scene(am3d_ctx* c, const am3d_scene* s) {
  API_BEGIN(c)
  validateScene(s);
  copyScene(c, s);
  resetState(c);
  c->haveScene = true;
  API_END(c)
}

int am3d_set_params(am3d_ctx* c, const am3d_params* p) {
  API_BEGIN(c)
  if (!p) throw AmError(AM3D_EINVAL, "null params");
  checkParams(p);
  c->P = *p;
  API_END(c)
}
int am3d_get_params(const am3d_ctx* c, am3d_params* p) {
  if (!c || !p) return AM3D_EINVAL;
  *p = c->P;
  return AM3D_OK;
}

int am3d_reset(am3d_ctx* c) {
  API_BEGIN(c)
  if (!c->haveScene) throw AmError(AM3D_ESTATE, "no scene uploaded");
  resetState(c);
  API_END(c)
}

int am3d_step(am3d_ctx* c, double dt, int nsteps) {
  API_BEGIN(c)
  if (!c->haveScene) throw AmError(AM3D_ESTATE, "no scene uploaded");
  for (int i = 0; i < nsteps; i++) stepOnce(c, dt);
  API_END(c)
}
int am3d_step_async(am3d_ctx* c, double dt, int nsteps) { return am3d_step(c, dt, nsteps); }
int am3d_sync(am3d_ctx* c) {
  API_BEGIN(c)
  CK(cudaStreamSynchronize(c->stream));
  API_END(c)
}

int am3d_num_bodies(const am3d_ctx* c) { return c ? c->NB : AM3D_EINVAL; }
int am3d_total_steps(const am3d_ctx* c) { return c ? c->totalSteps : AM3D_EINVAL; }

int am3d_download_bodies(am3d_ctx* c, double* x, double* R, double* v, double* omega, int32_t* sleeping, int32_t* collection) {
  API_BEGIN(c)
  if (!c->haveScene) throw AmError(AM3D_ESTATE, "no scene uploaded");
  int nb = c->NB;
  if (x) CK(cudaMemcpyAsync(x, c->x.p, 3 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (R) CK(cudaMemcpyAsync(R, c->R.p, 9 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (v) CK(cudaMemcpyAsync(v, c->v.p, 3 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (omega) CK(cudaMemcpyAsync(omega, c->w.p, 3 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  std::vector<int> fl(c->NS), par(nb);
  CK(cudaMemcpyAsync(fl.data(), c->flags.p, c->NS * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(par.data(), c->parent.p, nb * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < nb; i++) {
    int t = par[i] >= 0 ? par[i] : i;
    if (sleeping) sleeping[i] = (fl[t] & AM3D_F_SLEEPING) ? 1 : 0;
    if (collection) collection[i] = par[i] >= 0 ? par[i] - nb : -1;
  }
  API_END(c)
}

int am3d_upload_bodies(am3d_ctx* c, const double* x, const double* R, const double* v, const double* omega) {
  API_BEGIN(c)
  if (!c->haveScene) throw AmError(AM3D_ESTATE, "no scene uploaded");
  if (c->nCollections) throw AmError(AM3D_ESTATE, "cannot overwrite body state while collections exist");
  int nb = c->NB;
  CK(cudaMemcpyAsync(c->x.p, x, 3 * nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->R.p, R, 9 * nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->v.p, v, 3 * nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->w.p, omega, 3 * nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  // world-frame inertia follows the pose
  std::vector<double> jinv((size_t)nb * 9, 0.0), mA((size_t)nb * 9, 0.0);
  std::vector<int> fl(nb);
  CK(cudaMemcpyAsync(fl.data(), c->flags.p, nb * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(jinv.data(), c->jinv.p, jinv.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(mA.data(), c->mA.p, mA.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int i = 0; i < nb; i++) {
    if (fl[i] & AM3D_F_PINNED) continue;
    m3 Rm = ldm(R + 9 * i);
    stm(&jinv[9 * i], rm0rt(Rm, ldm(&c->H.body_jinv0[9 * i])));
    stm(&mA[9 * i], rm0rt(Rm, ldm(&c->H.body_mA0[9 * i])));
  }
  CK(cudaMemcpy(c->jinv.p, jinv.data(), jinv.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->mA.p, mA.data(), mA.size() * sizeof(double), cudaMemcpyHostToDevice));
  API_END(c)
}

int am3d_set_body_velocity(am3d_ctx* c, int body, const double v[3], const double omega[3]) {
  API_BEGIN(c)
  if (!c->haveScene || body < 0 || body >= c->NB) throw AmError(AM3D_EINVAL, "bad body index");
  if (v) CK(cudaMemcpyAsync(c->v.p + 3 * body, v, 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (omega) CK(cudaMemcpyAsync(c->w.p + 3 * body, omega, 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END(c)
}
int am3d_add_body_velocity(am3d_ctx* c, int body, const double dv[3], const double domega[3]) {
  API_BEGIN(c)
  if (!c->haveScene || body < 0 || body >= c->NB) throw AmError(AM3D_EINVAL, "bad body index");
  int par;
  CK(cudaMemcpy(&par, c->parent.p + body, sizeof(int), cudaMemcpyDeviceToHost));
  int t = par >= 0 ? par : body;
  double cur[3];
  if (dv) {
    CK(cudaMemcpy(cur, c->v.p + 3 * t, sizeof(cur), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) cur[k] = cur[k] + dv[k];
    CK(cudaMemcpy(c->v.p + 3 * t, cur, sizeof(cur), cudaMemcpyHostToDevice));
  }
  if (domega) {
    CK(cudaMemcpy(cur, c->w.p + 3 * t, sizeof(cur), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) cur[k] = cur[k] + domega[k];
    CK(cudaMemcpy(c->w.p + 3 * t, cur, sizeof(cur), cudaMemcpyHostToDevice));
  }
  API_END(c)
}

int am3d_num_contacts(am3d_ctx* c, int include_internal) {
  (void)include_internal;
  return c ? c->cur.n : AM3D_EINVAL;
}

int am3d_download_contacts(am3d_ctx* c, am3d_contact* out, int capacity, int include_internal, int* count) {
  API_BEGIN(c)
  (void)include_internal;
  int n = std::min(c->cur.n, capacity);
  if (count) *count = n;
  if (n > 0) {
    auto& S = c->cur;
    std::vector<int> b1(n), b2(n), s1(n), s2(n), bv1(n), bv2(n), info(n), leaf(n), state(n), isNew(n), bpc(n);
    std::vector<double> pW(3 * n), nW(3 * n), pB1(3 * n), nB1(3 * n), t1B1(3 * n), t2B1(3 * n), viol(n), pviol(n), lam(3 * n), lamW(3 * n);
    auto gi = [&](std::vector<int>& d, DevBuf<int>& s) { CK(cudaMemcpyAsync(d.data(), s.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream)); };
    auto gd = [&](std::vector<double>& d, DevBuf<double>& s, int w) { CK(cudaMemcpyAsync(d.data(), s.p, (size_t)w * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream)); };
    gi(b1, S.b1); gi(b2, S.b2); gi(s1, S.s1); gi(s2, S.s2); gi(bv1, S.bv1); gi(bv2, S.bv2); gi(info, S.info); gi(leaf, S.leaf);
    gi(state, S.state); gi(isNew, S.isNew); gi(bpc, S.bpc);
    gd(pW, S.pW, 3); gd(nW, S.nW, 3); gd(pB1, S.pB1, 3); gd(nB1, S.nB1, 3); gd(t1B1, S.t1B1, 3); gd(t2B1, S.t2B1, 3);
    gd(viol, S.viol, 1); gd(pviol, S.prevViol, 1); gd(lam, S.lam, 3); gd(lamW, S.lamWarm, 3);
    std::vector<int> gcol, gpos;
    int ng = c->nGroups;
    if (ng > 0 && ng == c->bp.n) {
      gcol.resize(ng); gpos.resize(ng);
      CK(cudaMemcpyAsync(gcol.data(), c->grpColor.p, ng * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaMemcpyAsync(gpos.data(), c->grpPos.p, ng * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) {
      am3d_contact& o = out[i];
      o.body1 = b1[i]; o.body2 = b2[i];
      o.csb1 = c->H.body_type[b1[i]] == AM3D_BODY_COMPOSITE ? s1[i] - c->H.body_shape_first[b1[i]] : -1;
      o.csb2 = c->H.body_type[b2[i]] == AM3D_BODY_COMPOSITE ? s2[i] - c->H.body_shape_first[b2[i]] : -1;
      o.bv1 = bv1[i]; o.bv2 = bv2[i]; o.info = info[i]; o.leaf = leaf[i]; o.state = state[i]; o.new_this_step = isNew[i];
      o.color = (!gcol.empty() && bpc[i] >= 0) ? gcol[bpc[i]] : -1;
      o.in_collection = 0;
      for (int k = 0; k < 3; k++) {
        o.contactB1[k] = pB1[3 * i + k]; o.normalB1[k] = nB1[3 * i + k]; o.tangent1B1[k] = t1B1[3 * i + k]; o.tangent2B1[k] = t2B1[3 * i + k];
        o.point_w[k] = pW[3 * i + k]; o.normal_w[k] = nW[3 * i + k]; o.lambda[k] = lam[3 * i + k]; o.lambda_warm[k] = lamW[3 * i + k];
      }
      o.violation = viol[i]; o.prev_violation = pviol[i];
    }
  }
  API_END(c)
}

int am3d_num_bpcs(am3d_ctx* c) { return c ? c->bpPrev.n : AM3D_EINVAL; }
int am3d_download_bpcs(am3d_ctx* c, am3d_bpc* out, int capacity, int* count) {
  API_BEGIN(c)
  int n = std::min(c->bpPrev.n, capacity);
  if (count) *count = n;
  if (n > 0) {
    std::vector<int> b1(n), b2(n), nm(n), ns(n), sh(4 * n);
    std::vector<double> mh(4 * n);
    CK(cudaMemcpyAsync(b1.data(), c->bpPrev.b1.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(b2.data(), c->bpPrev.b2.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(nm.data(), c->bpPrev.nMetric.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(ns.data(), c->bpPrev.nState.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(sh.data(), c->bpPrev.stateHist.p, 4 * n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(mh.data(), c->bpPrev.metricHist.p, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) {
      out[i].body1 = b1[i]; out[i].body2 = b2[i]; out[i].in_collection = 0; out[i].n_contacts = 0;
      out[i].n_metric = nm[i]; out[i].n_state = ns[i];
      for (int k = 0; k < 4; k++) { out[i].metric_hist[k] = mh[4 * i + k]; out[i].state_hist[k] = sh[4 * i + k]; }
    }
  }
  API_END(c)
}

int am3d_get_timings(am3d_ctx* c, am3d_timings* t) {
  if (!c || !t) return AM3D_EINVAL;
  *t = c->T;
  return AM3D_OK;
}

// ---- phase-level entry points --------------------------------------------------------------------
int am3d_detect(am3d_ctx* c) {
  API_BEGIN(c)
  if (!c->haveScene) throw AmError(AM3D_ESTATE, "no scene uploaded");
  detect(c);
  buildBodyPairs(c);
  CK(cudaStreamSynchronize(c->stream));
  API_END(c)
}

int am3d_solve(am3d_ctx* c, double dt) {
  API_BEGIN(c)
  if (!c->haveScene) throw AmError(AM3D_ESTATE, "no scene uploaded");
  applyExternalForces(c);  // clear + gravity + springs, deltaV = 0
  solveFull(c, dt, true);
  API_END(c)
}

int am3d_set_lambdas(am3d_ctx* c, const double* lam, int count) {
  API_BEGIN(c)
  if (count != c->cur.n) throw AmError(AM3D_EINVAL, "lambda count does not match the contact count");
  if (count) CK(cudaMemcpyAsync(c->cur.lam.p, lam, 3 * (size_t)count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END(c)
}

int am3d_download_deltav(am3d_ctx* c, double* dv) {
  API_BEGIN(c)
  CK(cudaMemcpyAsync(dv, c->dv.p, 6 * (size_t)c->NB * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END(c)
}

int am3d_download_solve_order(am3d_ctx* c, int32_t* order, int capacity, int* count) {
  API_BEGIN(c)
  int n = c->cur.n;
  if (count) *count = n;
  if (n > capacity) throw AmError(AM3D_EINVAL, "capacity too small");
  if (n > 0) {
    if (c->nGroups != c->bp.n || c->nGroups == 0) throw AmError(AM3D_ESTATE, "no solve has been run on the current contacts");
    // order[k] = canonical contact index solved k-th
    CK(cudaMemcpyAsync(order, c->scSrc.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  API_END(c)
}

int am3d_stats(am3d_ctx* c, double* out /* [4]: kernel launches, solve launches, row updates, solve kernel seconds */) {
  if (!c || !out) return AM3D_EINVAL;
  out[0] = (double)c->kernelLaunches; out[1] = (double)c->solveLaunches; out[2] = c->rowUpdates; out[3] = c->solveSeconds;
  return AM3D_OK;
}

int am3d_upload_contacts(am3d_ctx* c, const am3d_contact* in, int count) {
  (void)in; (void)count;
  if (!c) return AM3D_EINVAL;
  c->lastError = "am3d_upload_contacts is not built yet";
  return AM3D_EUNSUPPORTED;
}

}  // extern "C"
