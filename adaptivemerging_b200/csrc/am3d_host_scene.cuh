// Scene upload and reset: RigidBodySystem.reset() (RigidBodySystem.java:390-426) on the device arrays.
#pragma once
#include "am3d_host_util.cuh"
#include "am3d_sort.cuh"
#include "am3d_step.cuh"
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
  // collection slots: at most NB/2 collections are alive, and an unmerge can found up to NB/4 new ones before the
  // ones it dissolves are retired
  int NB = H.nb, NC = NB - NB / 4 + 2, NS = NB + NC;
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
  c->nDormant = 0;
  for (int i = 0; i < NB; i++) {
    fl[i] &= ~AM3D_F_SLEEPING;
    if (H.body_type[i] == AM3D_BODY_PLANE) fl[i] |= AM3D_F_PINNED;
    if (fl[i] & AM3D_F_DORMANT) { fl[i] |= AM3D_F_PINNED | AM3D_F_SLEEPING; c->nDormant++; }  // neither integrated nor collided
  }
  c->picked.ensure(NB + 1);
  CK(cudaMemsetAsync(c->picked.p, 0, (NB + 1) * sizeof(int), c->stream));
  {
    MouseState ms;
    memset(&ms, 0, sizeof(ms));
    ms.springBody = -1;
    c->mouse.ensure(sizeof(MouseState));
    CK(cudaMemcpyAsync(c->mouse.p, &ms, sizeof(ms), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->mouseUsed = false;
  }
  h2dv(c, c->flags, padI(fl, 0));
  h2dv(c, c->scene, padI(H.body_scene, 0));
  {  // scene-local body ids (rank of a body among the bodies of its scene)
    std::vector<int> cnt(H.nscenes, 0), loc(NB);
    for (int i = 0; i < NB; i++) loc[i] = cnt[H.body_scene[i]]++;
    h2dv(c, c->bodyLocal, loc);
  }
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
  c->force.ensure(3 * NS); c->torque.ensure(3 * NS); c->dv.ensure(DVS * (size_t)NS + 8);
  c->force.zero(3 * NS, c->stream); c->torque.zero(3 * NS, c->stream); c->dv.zero(DVS * (size_t)NS, c->stream);
  c->metricHist.ensure(10 * NS); c->metricHist.zero(10 * NS, c->stream);
  c->metricCount.ensure(NS); c->metricCount.zero(NS, c->stream);
  c->hasExt.ensure(NS); c->hasExt.zero(NS, c->stream);
  c->collAlive.ensure(NC); c->collAlive.zero(NC, c->stream);
  c->bodyBest.ensure(NS); c->bodyMask.ensure(NS);
  // merging state
  c->collMode.ensure(NC + 2); c->collMode.zero(NC + 2, c->stream);
  c->collFlagAcc.ensure(NC + 2); c->collFlagAcc.zero(NC + 2, c->stream);
  c->collCount.ensure(NC + 2); c->collCount.zero(NC + 2, c->stream);
  c->collStart.ensure(NC + 2); c->collStart.zero(NC + 2, c->stream);
  c->B2CR.ensure(9 * (size_t)NB); c->B2Ct.ensure(3 * (size_t)NB);
  c->B2CR.zero(9 * (size_t)NB, c->stream); c->B2Ct.zero(3 * (size_t)NB, c->stream);
  c->icon.n = 0; c->icon.nSorted = 0; c->ibp.n = 0; c->icon.ensure(256); c->ibp.ensure(64); c->ibpCut.ensure(64);
  c->events.clear(); c->orderFull.clear(); c->orderSweep.clear(); c->bpTail = false;
  // shapes
  h2dv(c, c->shType, H.shape_type); h2dv(c, c->shBody, H.shape_body); h2dv(c, c->shRoot, H.shape_root);
  h2dv(c, c->shSize, H.shape_size); h2dv(c, c->shRadius, H.shape_radius); h2dv(c, c->shP, H.shape_p);
  h2dv(c, c->shLR, H.shape_lR); h2dv(c, c->shLt, H.shape_lt);
  c->shX.ensure(3 * H.nsh); c->shR.ensure(9 * H.nsh); c->shBoundC.ensure(3 * H.nsh); c->shBoundR.ensure(H.nsh); c->shBoundH.ensure(3 * H.nsh);
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
  c->haveComposites = false;
  for (int b = 0; b < H.nb; b++) if (H.body_shape_count[b] > 1) c->haveComposites = true;
  c->hSmall.clear(); c->hLarge.clear(); c->hPlanes.clear();
  std::vector<int> isLarge(H.nsh, 0);
  for (int s = 0; s < H.nsh; s++) {
    if (H.shape_type[s] == AM3D_SHAPE_PLANE) c->hPlanes.push_back(s);
    else if (br[s] > thr) { c->hLarge.push_back(s); isLarge[s] = 1; }
    else { c->hSmall.push_back(s); maxSmall = std::max(maxSmall, br[s]); }
  }
  if (c->hLarge.size() > 4096 * (size_t)H.nscenes) {  // degenerate size distribution: fall back to one class
    for (int s : c->hLarge) { c->hSmall.push_back(s); isLarge[s] = 0; maxSmall = std::max(maxSmall, br[s]); }
    c->hLarge.clear();
    std::sort(c->hSmall.begin(), c->hSmall.end());
  }
  c->cellSize = maxSmall > 0 ? 2.0 * maxSmall * 1.0000001 : 1.0;
  // large shapes and planes grouped by scene: a shape only meets the special shapes of its own scene
  auto sceneOf = [&](int s) { return H.body_scene[H.shape_body[s]]; };
  auto byScene = [&](std::vector<int>& list, std::vector<int>& start) {
    std::stable_sort(list.begin(), list.end(), [&](int a, int b) { return sceneOf(a) < sceneOf(b); });
    start.assign(H.nscenes + 1, 0);
    for (int s : list) start[sceneOf(s) + 1]++;
    for (int k = 0; k < H.nscenes; k++) start[k + 1] += start[k];
  };
  std::vector<int> largeStart, planeStart;
  byScene(c->hLarge, largeStart);
  byScene(c->hPlanes, planeStart);
  h2dv(c, c->largeStart, largeStart); h2dv(c, c->planeStart, planeStart);
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
    c->nBodyBodySprings = 0;
    for (int s = 0; s < H.nsp; s++) c->nBodyBodySprings += H.sp_type[s] == AM3D_SPRING_BODYBODY;
    h2dv(c, c->spBodies, bodies); h2dv(c, c->spBodyStart, start); h2dv(c, c->spBodyList, list);
  }
  c->counters.ensure(64); c->counters.zero(64, c->stream);
  c->colorCtl.ensure(8);
  c->iterState.ensure(8); c->iterState.zero(8, c->stream);
  c->cur.n = 0; c->prev.n = 0; c->bp.n = 0; c->bpPrev.n = 0; c->cur.nSorted = 0; c->prev.nSorted = 0;
  {
    // reserve for ~12 contacts and ~5 body pairs per body up front so that steady growth of the contact count (a pile
    // settling layer by layer) does not reallocate every few steps
    size_t rc = (size_t)NB * 12 + 4096, rb = (size_t)NB * 5 + 1024;
    c->cur.ensure(rc); c->prev.ensure(rc); c->bp.ensure(rb); c->bpPrev.ensure(rb);
    c->hitPos.ensure(3 * rc * 2); c->hitNrm.ensure(3 * rc * 2); c->hitViol.ensure(rc * 2); c->hitMeta.ensure(4 * rc * 2);
    c->scP.ensure(24 * rc); c->hubDelta.ensure(12 * rb); c->grpDegree.ensure(NS + 1); c->grpHubMask.ensure(rb);
    c->scSrc.ensure(rc); c->scState.ensure(rc);
    c->sgB1.ensure(rb); c->sgB2.ensure(rb); c->sgStart.ensure(rb + 2); c->sgCount.ensure(rb + 2); c->sgFlags.ensure(rb);
    c->sgBpc.ensure(rb); c->sgMass.ensure(20 * rb); c->sgMu.ensure(rb);
    c->tmpI0.ensure(rc); c->tmpI1.ensure(rc); c->tmpI2.ensure(rb); c->tmpI3.ensure(rb);
    if (c->P.enable_merging) {  // internal tables of the collections (and their compaction targets)
      c->icon.ensure(rc / 2); c->icon2.ensure(rc / 2); c->ibp.ensure(rb / 2); c->ibp2.ensure(rb / 2);
      c->ibpCut.ensure(rb / 2); c->ibpCut2.ensure(rb / 2);
    }
  }
  c->totalSteps = 0;
  c->mergingEvent = false;
  c->nCollections = 0;
  c->nMergedLeaves = 0;
  c->nextStamp = NS;
  memset(&c->T, 0, sizeof(c->T));
  CK(cudaStreamSynchronize(c->stream));
}

// ------------------------------------------------------------------------------------------------
// detection
// ------------------------------------------------------------------------------------------------
