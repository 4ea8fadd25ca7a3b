// C ABI (include/am3d.h) and the per-step kernel sequence.
// Sequence = RigidBodySystem.advanceTime (RigidBodySystem.java:102-185); see stepOnce().
#include "am3d_host_util.cuh"
#include "am3d_sort.cuh"
#include "am3d_host_scene.cuh"
#include "am3d_host_detect.cuh"
#include "am3d_host_solve.cuh"

// ------------------------------------------------------------------------------------------------
// the step
// ------------------------------------------------------------------------------------------------
static void applyCoriolis(am3d_ctx* c, double dt, int topOnly) {
  if (c->P.use_coriolis)
    LAUNCH(c, k_coriolis, nblk(c->NS), BLK, c->NS, c->NB, c->collAlive.p, c->parent.p, c->flags.p, c->w.p, c->jinv.p, c->mA.p, c->torque.p, dt, topOnly);
}
static void applyMouseTools(am3d_ctx* c, int part) {  // RigidBodySystem.java:244-247 (part 0), :249-256 (part 1)
  if (c->mouseUsed)
    LAUNCH(c, k_mouse_tools, 1, 32, part, (MouseState*)c->mouse.p, c->parent.p, c->flags.p, c->metricCount.p, c->picked.p, c->x.p, c->R.p, c->v.p,
           c->w.p, c->force.p, c->torque.p);
}
static void applyExternalForces(am3d_ctx* c, double dt) {
  const am3d_params& P = c->P;
  double theta = P.gravity_angle_deg / 180.0 * M_PI;
  double gx = P.gravity_amount * cos(theta), gy = P.gravity_amount * sin(theta);
  LAUNCH(c, k_clear_gravity, nblk(c->NS), BLK, c->NS, c->NB, c->collAlive.p, c->parent.p, c->mass.p, c->x.p, c->v.p, c->w.p,
         c->force.p, c->torque.p, c->dv.p, P.use_gravity, gx, gy);
  applyCoriolis(c, dt, 0);
  applyMouseTools(c, 0);
  if (P.springs_enabled && c->nSpringBodies > 0)
    LAUNCH(c, k_springs, nblk(c->nSpringBodies, 64), 64, c->nSpringBodies, c->spBodies.p, c->spBodyStart.p, c->spBodyList.p,
           c->spType.p, c->spB1.p, c->spB2.p, c->spPb1.p, c->spPb2.p, c->spPw.p, c->spK.p, c->spD.p, c->spL0.p, c->spLs.p,
           P.spring_k_mod, P.spring_d_mod, c->parent.p, c->x.p, c->R.p, c->v.p, c->w.p, c->force.p, c->torque.p);
  applyMouseTools(c, 1);
}

static void applySprings(am3d_ctx* c) {
  const am3d_params& P = c->P;
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

// fill `out[0..n)` from the first n entries of a contact set
static void fetchContacts(am3d_ctx* c, ContactSet& S, int n, am3d_contact* out, int internal) {
  if (n <= 0) return;
  std::vector<int> b1(n), b2(n), s1(n), s2(n), bv1(n), bv2(n), info(n), leaf(n), state(n), isNew(n), bpc(n);
  std::vector<double> pW(3 * n), nW(3 * n), pB1(3 * n), nB1(3 * n), t1B1(3 * n), t2B1(3 * n), viol(n), pviol(n), lam(3 * n), lamW(3 * n);
  auto gi = [&](std::vector<int>& d, DevBuf<int>& s) { CK(cudaMemcpyAsync(d.data(), s.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream)); };
  auto gd = [&](std::vector<double>& d, DevBuf<double>& s, int w) { CK(cudaMemcpyAsync(d.data(), s.p, (size_t)w * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream)); };
  gi(b1, S.b1); gi(b2, S.b2); gi(s1, S.s1); gi(s2, S.s2); gi(bv1, S.bv1); gi(bv2, S.bv2); gi(info, S.info); gi(leaf, S.leaf);
  gi(state, S.state); gi(isNew, S.isNew); gi(bpc, S.bpc);
  gd(pW, S.pW, 3); gd(nW, S.nW, 3); gd(pB1, S.pB1, 3); gd(nB1, S.nB1, 3); gd(t1B1, S.t1B1, 3); gd(t2B1, S.t2B1, 3);
  gd(viol, S.viol, 1); gd(pviol, S.prevViol, 1); gd(lam, S.lam, 3); gd(lamW, S.lamWarm, 3);
  std::vector<int> gcol;
  int ng = c->nGroups;
  if (!internal && !c->lastSolveSweep && ng > 0 && ng == c->bp.n && c->nPairsSolve == ng) {
    gcol.resize(ng);
    CK(cudaMemcpyAsync(gcol.data(), c->grpColor.p, ng * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; i++) {
    am3d_contact& o = out[i];
    o.body1 = b1[i]; o.body2 = b2[i];
    o.csb1 = c->H.body_type[b1[i]] == AM3D_BODY_COMPOSITE ? s1[i] - c->H.body_shape_first[b1[i]] : -1;
    o.csb2 = c->H.body_type[b2[i]] == AM3D_BODY_COMPOSITE ? s2[i] - c->H.body_shape_first[b2[i]] : -1;
    o.bv1 = bv1[i]; o.bv2 = bv2[i]; o.info = info[i]; o.leaf = leaf[i]; o.state = state[i]; o.new_this_step = isNew[i];
    o.color = (!gcol.empty() && bpc[i] >= 0 && bpc[i] < ng) ? gcol[bpc[i]] : -1;
    o.in_collection = internal;
    o.hub_mask = 0; o._pad = 0;
    for (int k = 0; k < 3; k++) {
      o.contactB1[k] = pB1[3 * i + k]; o.normalB1[k] = nB1[3 * i + k]; o.tangent1B1[k] = t1B1[3 * i + k]; o.tangent2B1[k] = t2B1[3 * i + k];
      o.point_w[k] = pW[3 * i + k]; o.normal_w[k] = nW[3 * i + k]; o.lambda[k] = lam[3 * i + k]; o.lambda_warm[k] = lamW[3 * i + k];
    }
    o.violation = viol[i]; o.prev_violation = pviol[i];
  }
}

// tests: turn the recorded solve sequence into contact identities while the tables it indexes are still intact
static void snapshotOrder(am3d_ctx* c, int which) {
  std::vector<int>& ord = which == 1 ? c->orderSweep : which == 2 ? c->orderPost : c->orderFull;
  std::vector<am3d_contact>& dst = which == 1 ? c->orderSweepKeys : which == 2 ? c->orderPostKeys : c->orderFullKeys;
  dst.clear();
  CK(cudaStreamSynchronize(c->stream));
  int n = (int)ord.size();
  if (n == 0) return;
  std::vector<am3d_contact> ext(c->cur.n), in(which == 1 ? c->icon.n : 0);
  fetchContacts(c, c->cur, c->cur.n, ext.data(), 0);
  if (which == 1) fetchContacts(c, c->icon, c->icon.n, in.data(), 1);
  // per solve position: dense colour index and hub sides of the group the contact belongs to
  int ng = c->nGroups;
  std::vector<int> sgStart(ng), sgCount(ng), sgFlags(ng), sgPhase(ng);
  CK(cudaMemcpy(sgStart.data(), c->sgStart.p, ng * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sgCount.data(), c->sgCount.p, ng * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sgFlags.data(), c->sgFlags.p, ng * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sgPhase.data(), c->sgPhase.p, ng * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<int> posColor(n, 0), posHub(n, 0);
  for (int p = 0; p < ng; p++)
    for (int k = 0; k < sgCount[p]; k++) {
      int idx = sgStart[p] + k;
      if (idx < n) { posColor[idx] = sgPhase[p]; posHub[idx] = ((sgFlags[p] & SG_HUB1) ? 1 : 0) | ((sgFlags[p] & SG_HUB2) ? 2 : 0); }
    }
  dst.resize(n);
  for (int k = 0; k < n; k++) {
    int set = ord[k] >> 30, i = ord[k] & 0x3fffffff;
    dst[k] = set ? in[i] : ext[i];
    dst[k].color = posColor[k];
    dst[k].hub_mask = posHub[k];
  }
}

// sort next step's body-pair lookup table by key after an unmerge appended pairs out of order
static void sortBpPrev(am3d_ctx* c) {
  int n = c->bpPrev.n;
  if (n < 2) return;
  c->tmpI0.ensure(n + 2); c->tmpI1.ensure(n + 2); c->grpKey.ensure(n + 2); c->grpKeySorted.ensure(n + 2);
  LAUNCH(c, k_iota, nblk(n), BLK, n, c->tmpI0.p);
  sortPairs(c, c->bpPrev.key.p, c->grpKeySorted.p, c->tmpI0.p, c->tmpI1.p, n, 0, 48);
  c->bpTmp.ensure(n + 1);
  LAUNCH(c, k_bpc_gather, nblk(n), BLK, n, c->tmpI1.p, c->bpPrev.key.p, c->bpPrev.b1.p, c->bpPrev.b2.p, c->bpPrev.metricHist.p,
         c->bpPrev.stateHist.p, c->bpPrev.nMetric.p, c->bpPrev.nState.p, c->bpTmp.key.p, c->bpTmp.b1.p, c->bpTmp.b2.p,
         c->bpTmp.metricHist.p, c->bpTmp.stateHist.p, c->bpTmp.nMetric.p, c->bpTmp.nState.p);
  std::swap(c->bpPrev, c->bpTmp);
  c->bpPrev.n = n;
}

// surviving external body pairs become the lookup table of the next detection (histories carry over)
static void carryBodyPairs(am3d_ctx* c) {
  int nbp = c->bp.n, nAlive = 0;
  if (nbp > 0) {
    c->tmpI1.ensure(nbp + 2);
    nAlive = scanTotal(c, c->bp.alive, c->tmpI1, nbp);
    c->bpPrev.ensure(nAlive + 1);
    LAUNCH(c, k_bpc_compact, nblk(nbp), BLK, nbp, c->bp.alive.p, c->tmpI1.p, c->bp.key.p, c->bp.b1.p, c->bp.b2.p, c->bp.metricHist.p,
           c->bp.stateHist.p, c->bp.nMetric.p, c->bp.nState.p, c->bpPrev.key.p, c->bpPrev.b1.p, c->bpPrev.b2.p,
           c->bpPrev.metricHist.p, c->bpPrev.stateHist.p, c->bpPrev.nMetric.p, c->bpPrev.nState.p);
  }
  c->bpPrev.n = nAlive;
  if (c->bpTail) sortBpPrev(c);
}

// RigidBodySystem.postStabilization (:354-377): detect again at the advanced positions, solve the position-level problem
// (right-hand side = feedbackStiffness * violation, PGS.java:86-89) and move the bodies by the resulting deltaV
static void postStabilization(am3d_ctx* c, double dt) {
  int NS = c->NS, NB = c->NB;
  carryBodyPairs(c);  // the BodyPairContact objects (with this step's histories) are what the second detection refills
  if (c->nCollections > 0)  // RigidCollection.clearBodies :100-105: members take the collection's velocity at their new positions
    LAUNCH(c, k_members_take_velocity, nblk(NB), BLK, NB, c->parent.p, c->x.p, c->v.p, c->w.p);
  CK(cudaMemsetAsync(c->force.p, 0, 3 * (size_t)NS * sizeof(double), c->stream));   // RigidBody.clear :276-280
  CK(cudaMemsetAsync(c->torque.p, 0, 3 * (size_t)NS * sizeof(double), c->stream));
  CK(cudaMemsetAsync(c->dv.p, 0, DVS * (size_t)NS * sizeof(double), c->stream));
  detect(c);
  buildBodyPairs(c);
  warmStart(c, true);
  runSolve(c, dt, false, true);
  c->orderPostKeys.clear();
  if (c->recordOrders) snapshotOrder(c, 2);
  int nbp = c->bp.n;
  if (nbp > 0) {
    if (c->cur.n == 0) CK(cudaMemsetAsync(c->bp.nActive.p, 0, nbp * sizeof(int), c->stream));
    LAUNCH(c, k_bpc_prune, nblk(nbp), BLK, nbp, c->bp.nActive.p, c->bp.alive.p);  // clearBodyPairContacts :374
  }
  // advancePositionsPostStabilization (RigidBody.java:423-425): advancePositions with (deltaV.v, deltaV.w)
  LAUNCH(c, k_advance_positions, nblk(NS), BLK, NS, NB, c->collAlive.p, c->parent.p, c->flags.p, c->x.p, c->R.p, c->dv.p, c->dv.p + 3, DVS,
         c->jinv0.p, c->mA0.p, c->jinv.p, c->mA.p, dt);
  if (c->nCollections > 0)
    LAUNCH(c, k_members_follow, nblk(NB), BLK, NB, c->parent.p, c->flags.p, 0, 1, c->x.p, c->R.p, c->v.p, c->w.p, c->B2CR.p, c->B2Ct.p,
           c->jinv0.p, c->mA0.p, c->jinv.p, c->mA.p);
}

// a velocity poke uploaded by am3d_add_velocities and not yet added to the velocities
static void flushPokes(am3d_ctx* c) {
  if (!c->pokesPending) return;
  CK(cudaStreamWaitEvent(c->stream, c->evPoke, 0));
  LAUNCH(c, k_add_velocities, nblk(c->NB), BLK, c->NB, c->parent.p, c->pokeV.p, c->pokeW.p, c->v.p, c->w.p);
  c->pokesPending = false;
}
static void stepOnce(am3d_ctx* c, double dt) {
  const am3d_params& P = c->P;
  int NS = c->NS, NB = c->NB;
  c->totalSteps++;
  double theta = P.gravity_angle_deg / 180.0 * M_PI;
  double gx = P.gravity_amount * cos(theta), gy = P.gravity_amount * sin(theta);
  CK(cudaEventRecord(c->ev[0], c->stream));
  // a velocity poke still on its way up (am3d_add_velocities): detection reads positions only, so it goes first and the
  // poke + the external forces (springs read velocities) follow it - the same values as in the reference's order
  const bool pokeLate = c->pokesPending;
  if (!pokeLate) applyExternalForces(c, dt);    // clearBodies + applyExternalForces  (:108-116)
  CK(cudaEventRecord(c->ev[1], c->stream));
  detect(c);                                    // updateContactsMap + collisionDetection (:119-120)
  buildBodyPairs(c);                            // updateBodyPairContacts (:121)
  CK(cudaEventRecord(c->ev[2], c->stream));
  if (pokeLate) { flushPokes(c); applyExternalForces(c, dt); }
  warmStart(c);                                 // warmStart(false) (:124)
  CK(cudaEventRecord(c->ev[3], c->stream));
  if (P.enable_sleeping) {                      // sleeping.wake() (:128)
    LAUNCH(c, k_wake_pairs, nblk(c->bp.n), BLK, c->bp.n, c->bp.b1.p, c->bp.b2.p, c->parent.p, c->flags.p, c->metricCount.p);
    if (c->nBodyBodySprings > 0)
      LAUNCH(c, k_wake_springs, nblk(c->NSP), BLK, c->NSP, c->spType.p, c->spB1.p, c->spB2.p, c->parent.p, c->flags.p, c->metricCount.p);
  }
  // single sweep over external + internal contacts (:131), only while collections exist
  c->T.update_collections = c->T.contact_ordering = c->T.single_it_pgs = 0;
  c->orderSweep.clear();
  c->orderSweepKeys.clear();
  bool swept = false;
  if (c->nCollections > 0 && P.update_contacts_in_collections) {
    runSolve(c, dt, true);
    LAUNCH(c, k_sweep_finish, nblk(NS), BLK, NS, NB, c->parent.p, c->flags.p, c->minv.p, c->jinv.p, c->force.p, c->torque.p, c->dv.p,
           c->v.p, c->w.p, dt);
    swept = true;
    if (c->recordOrders) snapshotOrder(c, 1);
  }
  CK(cudaEventRecord(c->ev[14], c->stream));
  // accumulateForUnmerging + unmerge (:135-136)
  if (c->nCollections > 0) unmergeStep(c, dt);
  if (c->mergingEvent) {                        // sticky flag (:138-142): clear + re-apply forces on top-level bodies
    LAUNCH(c, k_reclear_top, nblk(NS), BLK, NS, NB, c->collAlive.p, c->parent.p, c->mass.p, c->force.p, c->torque.p, c->dv.p,
           P.use_gravity, gx, gy);
    applyCoriolis(c, dt, 0);  // applyExternalForces runs again in full: second gyroscopic term, Coriolis torque on members too
    applyMouseTools(c, 0);
    applySprings(c);
    applyMouseTools(c, 1);
  }
  CK(cudaEventRecord(c->ev[4], c->stream));
  // redoWarmStart (:146-150): the sweep never writes the multipliers of external contacts back, so lambda == lambdaWarm
  if (!P.warm_start && c->cur.n)                // noWarmStart (:464-470)
    CK(cudaMemsetAsync(c->cur.lam.p, 0, 3 * (size_t)c->cur.n * sizeof(double), c->stream));
  runSolve(c, dt, false);                       // solveLCP (:151)
  c->orderFullKeys.clear();
  if (c->recordOrders) snapshotOrder(c, 0);
  CK(cudaEventRecord(c->ev[5], c->stream));
  // clearBodyPairContacts (:152) is folded into k_post_solve (nActive) + k_bpc_accumulate
  LAUNCH(c, k_advance_velocities, nblk(NS), BLK, NS, NB, c->collAlive.p, c->parent.p, c->flags.p, c->minv.p, c->jinv.p,
         c->force.p, c->torque.p, c->dv.p, c->v.p, c->w.p, dt);                                  // (:155)
  if (c->nCollections > 0)
    LAUNCH(c, k_members_follow, nblk(NB), BLK, NB, c->parent.p, c->flags.p, 1, 0, c->x.p, c->R.p, c->v.p, c->w.p, c->B2CR.p, c->B2Ct.p,
           c->jinv0.p, c->mA0.p, c->jinv.p, c->mA.p);
  CK(cudaMemsetAsync(c->hasExt.p, 0, NS * sizeof(int), c->stream));
  int nbp = c->bp.n;
  if (nbp > 0) {
    if (c->cur.n == 0) CK(cudaMemsetAsync(c->bp.nActive.p, 0, nbp * sizeof(int), c->stream));
    LAUNCH(c, k_bpc_accumulate, nblk(nbp, 128), 128, nbp, c->bp.b1.p, c->bp.b2.p, c->bp.start.p, c->bp.count.p, c->bp.nActive.p,
           c->bp.alive.p, c->cur.lam.p, c->cur.state.p, c->parent.p, c->flags.p, c->x.p, c->R.p, c->v.p, c->w.p, c->bbB.p,
           c->bbCount.p, c->bp.metricHist.p, c->bp.stateHist.p, c->bp.nMetric.p, c->bp.nState.p, P.step_accum_merging,
           c->hasExt.p, P.metric_position_level ? dt : 0.0);                                     // (:158)
  }
  LAUNCH(c, k_advance_positions, nblk(NS), BLK, NS, NB, c->collAlive.p, c->parent.p, c->flags.p, c->x.p, c->R.p, c->v.p,
         c->w.p, 3, c->jinv0.p, c->mA0.p, c->jinv.p, c->mA.p, dt);                                // (:160)
  if (c->nCollections > 0)
    LAUNCH(c, k_members_follow, nblk(NB), BLK, NB, c->parent.p, c->flags.p, 0, 1, c->x.p, c->R.p, c->v.p, c->w.p, c->B2CR.p, c->B2Ct.p,
           c->jinv0.p, c->mA0.p, c->jinv.p, c->mA.p);
  c->orderPostKeys.clear();
  if (P.enable_post_stabilization) { postStabilization(c, dt); nbp = c->bp.n; }                  // (:162-163)
  CK(cudaEventRecord(c->ev[6], c->stream));
  if ((c->totalSteps % P.steps_between_merge) == 0) mergeStep(c);                                // (:166-167)
  CK(cudaEventRecord(c->ev[15], c->stream));
  if (nbp > 0)
    LAUNCH(c, k_has_ext, nblk(nbp), BLK, nbp, c->bp.alive.p, c->bp.b1.p, c->bp.b2.p, c->parent.p, c->flags.p, c->hasExt.p);
  carryBodyPairs(c);
  if (P.enable_sleeping)                                                                         // (:170)
    LAUNCH(c, k_sleep, nblk(NS), BLK, NS, NB, c->collAlive.p, c->parent.p, c->flags.p, c->hasExt.p, c->metricHist.p,
           c->metricCount.p, c->x.p, c->R.p, c->v.p, c->w.p, c->bbB.p, c->bbCount.p, P.sleep_step_accum, P.sleep_threshold);
  LAUNCH(c, k_viscous, nblk(NS), BLK, NS, NB, c->collAlive.p, c->parent.p, c->v.p, c->w.p, P.viscous_linear, P.viscous_angular);  // (:173)
  CK(cudaEventRecord(c->ev[7], c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->T.detection = evMs(c, 1, 2) * 1e-3;
  c->T.narrowphase_kernel_time = c->narrowTimed ? (evMs(c, 16, 17) + evMs(c, 18, 19)) * 1e-3 : 0.0;
  c->narrowTimed = false;
  c->T.warmstart = evMs(c, 2, 3) * 1e-3;
  c->T.update_collections = swept ? evMs(c, 3, 14) * 1e-3 : 0;
  c->T.single_it_pgs = swept ? evMs(c, 8, 9) * 1e-3 : 0;
  c->T.contact_ordering = (swept && c->orderingTimed) ? evMs(c, 20, 21) * 1e-3 : 0;
  c->orderingTimed = false;
  c->T.unmerging = evMs(c, 14, 4) * 1e-3;
  c->T.lcp_solve = evMs(c, 4, 5) * 1e-3;
  c->T.merging = evMs(c, 6, 15) * 1e-3;
  // Merging.params.mergingBuildTime / unmergingBuildTime: the part after the merge / unmerge conditions were evaluated
  c->T.merging_build = c->mergeBuildTimed ? evMs(c, 22, 15) * 1e-3 : 0.0;
  c->T.unmerging_build = c->unmergeBuildTimed ? evMs(c, 23, 4) * 1e-3 : 0.0;
  c->mergeBuildTimed = c->unmergeBuildTimed = false;
  c->T.compute_time = evMs(c, 0, 7) * 1e-3;
  c->T.n_bodies = NB - c->nDormant - c->nMergedLeaves + c->nCollections;  // bodies.size(): a collection counts as one (RigidBodySystem.java:503)
  c->T.n_contacts = c->cur.n;  // collision.contacts.size() at the end of the step, unmerge-appended contacts included
  c->T.n_collections = c->nCollections;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
static void waitPending(am3d_ctx* c) {
  if (c->pending.valid()) {
    int rc = c->pending.get();
    if (rc != AM3D_OK && c->asyncStatus == AM3D_OK) c->asyncStatus = rc;
  }
}
// (_NOFLUSH: the entry points that leave an uploaded velocity poke pending, see am3d_add_velocities)
#define API_BEGIN_NOFLUSH(ctx)              \
  if (!(ctx)) return AM3D_EINVAL;           \
  try {                                     \
    waitPending(ctx);                       \
    cudaSetDevice((ctx)->device);           \
    amCurrentStream() = (ctx)->stream;      \
    amCurrentPool() = (ctx)->pool;
#define API_BEGIN(ctx)                      \
  API_BEGIN_NOFLUSH(ctx)                    \
    if ((ctx)->pokesPending) flushPokes(ctx);
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
  // shuffle (Collections.shuffle with an unseeded Random, CollisionProcessor.java:674): asks for an unspecified order of the
  // Gauss-Seidel sweep; the colour order already is one (and stays deterministic), so the flag is accepted as it is
  // collectionCD (CollisionProcessor.java:768-790): 0 = brute force over the members, 1 = the collection's sphere BVH
  // (:859-912, RigidCollection.java:184-453), 2 = world-AABB test per member pair (:921-960).  Modes 1 and 2 only PRUNE the
  // narrowPhase(member, other) calls of mode 0 with bounding volumes that enclose the members (BVSphere(body) encloses the
  // body's bounding box, getEnclosingSphere :414-453 encloses both children, sweepCollide compares the world boxes), so
  // all three modes yield the same contacts; they differ in the order contacts enter the list, which this library replaces by
  // its canonical order anyway.  Detection here always prunes leaf-body pairs by world AABB, whatever the mode.
  if (p->collection_cd < 0 || p->collection_cd > 2) throw AmError(AM3D_EINVAL, "collection_cd must be 0 (brute force), 1 (BVH) or 2 (sweep and prune)");
  if (p->merge_cycle_condition) throw AmError(AM3D_EUNSUPPORTED, "cycle merge condition is not supported");
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
    amCurrentStream() = c->stream;
    {  // a private pool that keeps freed blocks instead of returning them to the driver at every synchronisation
      cudaMemPoolProps props;
      memset(&props, 0, sizeof(props));
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = device;
      CK(cudaMemPoolCreate(&c->pool, &props));
      unsigned long long keep = ~0ULL;
      CK(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep));
      amCurrentPool() = c->pool;
    }
    for (int i = 0; i < 24; i++) CK(cudaEventCreate(&c->ev[i]));
    c->evCreated = true;
    am3d_default_params(&c->P);
    int coop = 0, sms = 0, perSm = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    // co-resident CTAs of the cooperative kernel, per variant [hub support]
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_pgs_persistent<false>, 128, 0)); c->coopBlocksV[0] = coop ? sms * perSm : 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_pgs_persistent<true>, 128, 0));  c->coopBlocksV[1] = coop ? sms * perSm : 0;
    c->coopBlocks = c->coopBlocksV[1];
    {
      int gsm = GIANT_WARPS * 32 * GIANT_ROW * (int)sizeof(double);
      CK(cudaFuncSetAttribute(k_pgs_giant<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gsm));
      CK(cudaFuncSetAttribute(k_pgs_giant<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gsm));
      CK(cudaFuncSetAttribute(k_pgs_giant<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gsm));
      CK(cudaFuncSetAttribute(k_pgs_giant<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gsm));
    }
    {  // how many 8-CTA clusters of the partitioned sweep fit the GPU at once
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(PGS_CLUSTER * sms);
      cfg.blockDim = dim3(128);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = PGS_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n1 = 0, n2 = 0;
      if (cudaOccupancyMaxActiveClusters(&n1, k_pgs_cluster<false>, &cfg) != cudaSuccess) { n1 = 0; cudaGetLastError(); }
      if (cudaOccupancyMaxActiveClusters(&n2, k_pgs_cluster<true>, &cfg) != cudaSuccess) { n2 = 0; cudaGetLastError(); }
      c->maxClusters = std::min(n1, n2);
      if (amTrace()) fprintf(stderr, "[am3d] co-resident clusters of %d CTAs: %d / %d\n", PGS_CLUSTER, n1, n2);
      if (const char* e = getenv("AM3D_PGS_CLUSTERS")) c->useClusters = atoi(e);
    }
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_bfs_layers, 256, 0)); c->bfsBlocks = coop ? sms * perSm : 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_color_coop, 256, 0)); c->colorBlocks = coop ? sms * perSm : 0;
    if (const char* e = getenv("AM3D_PGS_PERSISTENT")) c->usePersistent = atoi(e);
    if (const char* e = getenv("AM3D_GIANT_WARPS")) c->useGiantWarps = atoi(e);
    if (const char* e = getenv("AM3D_GIANT_CHUNK")) c->giantChunk = atoi(e);
  } catch (const AmError& e) {
    cudaMemPool_t pool = c->pool;
    delete c;
    if (pool) cudaMemPoolDestroy(pool);
    return e.code;
  }
  *out = c;
  return AM3D_OK;
}

int am3d_destroy(am3d_ctx* c) {
  if (!c) return AM3D_EINVAL;
  waitPending(c);
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->evCreated) for (int i = 0; i < 24; i++) cudaEventDestroy(c->ev[i]);
  if (c->mappedHost) cudaFreeHost(c->mappedHost);
  if (c->copyStream) {
    if (c->upStream) { cudaStreamSynchronize(c->upStream); cudaStreamDestroy(c->upStream); }
    if (c->evPoke) cudaEventDestroy(c->evPoke);
    if (c->evMain) cudaEventDestroy(c->evMain);
    cudaStreamSynchronize(c->copyStream);
    cudaEventDestroy(c->evSnap); cudaEventDestroy(c->evCopied);
    cudaStreamDestroy(c->copyStream);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  cudaMemPool_t pool = c->pool;
  delete c;  // frees every device buffer
  if (pool) cudaMemPoolDestroy(pool);
  if (amCurrentPool() == pool) amCurrentPool() = nullptr;
  return AM3D_OK;
}

const char* am3d_last_error(const am3d_ctx* c) { return c ? c->lastError.c_str() : "null context"; }

int am3d_upload_scene(am3d_ctx* c, const am3d_scene* s) {
  API_BEGIN(c)
  validateScene(s);
  size_t before = amAllocatedBytes();
  copyScene(c, s);
  resetState(c);
  {  // pre-warm the allocator pool: buffers regrow during a run (contacts and candidate pairs vary), and a pool that has to
     // go back to the driver for memory stalls the step by ~100 ms
    size_t warm = std::max<size_t>((amAllocatedBytes() - before) / 2, 64u << 20);
    void* tmp = nullptr;
    if (cudaMallocFromPoolAsync(&tmp, warm, c->pool, c->stream) == cudaSuccess) cudaFreeAsync(tmp, c->stream);
    else cudaGetLastError();
    CK(cudaStreamSynchronize(c->stream));
  }
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
  API_BEGIN_NOFLUSH(c)
  if (!c->haveScene) throw AmError(AM3D_ESTATE, "no scene uploaded");
  for (int i = 0; i < nsteps; i++) stepOnce(c, dt);
  API_END(c)
}
static int stepBody(am3d_ctx* c, double dt, int nsteps) {
  try {
    cudaSetDevice(c->device);
    amCurrentStream() = c->stream;
    amCurrentPool() = c->pool;
    for (int i = 0; i < nsteps; i++) stepOnce(c, dt);
    return AM3D_OK;
  } catch (const AmError& e) {
    c->lastError = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    c->lastError = e.what();
    return AM3D_ECUDA;
  }
}
int am3d_step_async(am3d_ctx* c, double dt, int nsteps) {
  API_BEGIN_NOFLUSH(c)  // waits for steps still pending from an earlier call
  if (!c->haveScene) throw AmError(AM3D_ESTATE, "no scene uploaded");
  if (c->asyncStatus != AM3D_OK) { int rc = c->asyncStatus; c->asyncStatus = AM3D_OK; return rc; }
  c->pending = std::async(std::launch::async, stepBody, c, dt, nsteps);
  API_END(c)
}
int am3d_sync(am3d_ctx* c) {
  API_BEGIN(c)
  CK(cudaStreamSynchronize(c->stream));
  if (c->asyncStatus != AM3D_OK) { int rc = c->asyncStatus; c->asyncStatus = AM3D_OK; return rc; }
  API_END(c)
}

int am3d_num_bodies(const am3d_ctx* c) { return c ? c->NB : AM3D_EINVAL; }
int am3d_total_steps(const am3d_ctx* c) { if (c) waitPending(const_cast<am3d_ctx*>(c)); return c ? c->totalSteps : AM3D_EINVAL; }

// Outbound read of the body state (what Display reads after a step).  The state is snapshot into staging buffers on
// the step's stream and copied to the host on a second stream, so that with the _async form the copy of step N runs
// under the kernels of step N+1; am3d_wait_download() (or the next download) waits for it.
__global__ void k_body_flags_out(int nb, const int* __restrict__ parent, const int* __restrict__ flags, int* __restrict__ sleeping,
                                 int* __restrict__ collection) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  int p = parent[i];
  int t = p >= 0 ? p : i;
  sleeping[i] = ((flags[t] & AM3D_F_SLEEPING) && !(flags[t] & AM3D_F_DORMANT)) ? 1 : 0;  // (dormant bodies carry the flag only to stay out of the integrator)
  collection[i] = p >= 0 ? p - nb : -1;
}
static void ensureCopyStream(am3d_ctx* c) {
  if (!c->copyStream) {
    CK(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->evSnap, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evCopied, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&c->upStream, cudaStreamNonBlocking));  // (uploads on their own stream: the two directions use different copy engines)
    CK(cudaEventCreateWithFlags(&c->evPoke, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evMain, cudaEventDisableTiming));
  }
}
static void downloadBodies(am3d_ctx* c, double* x, double* R, double* v, double* omega, int32_t* sleeping, int32_t* collection) {
  if (!c->haveScene) throw AmError(AM3D_ESTATE, "no scene uploaded");
  size_t nb = c->NB;
  ensureCopyStream(c);
  if (c->copyPending) { CK(cudaEventSynchronize(c->evCopied)); c->copyPending = false; }  // staging is free again
  c->stD.ensure(18 * nb + 8); c->stI.ensure(2 * nb + 8);
  double *sx = c->stD.p, *sR = sx + 3 * nb, *sv = sR + 9 * nb, *sw = sv + 3 * nb;
  int *ssl = c->stI.p, *sco = ssl + nb;
  if (x) CK(cudaMemcpyAsync(sx, c->x.p, 3 * nb * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (R) CK(cudaMemcpyAsync(sR, c->R.p, 9 * nb * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (v) CK(cudaMemcpyAsync(sv, c->v.p, 3 * nb * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (omega) CK(cudaMemcpyAsync(sw, c->w.p, 3 * nb * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (sleeping || collection) LAUNCH(c, k_body_flags_out, nblk((int)nb), BLK, (int)nb, c->parent.p, c->flags.p, ssl, sco);
  CK(cudaEventRecord(c->evSnap, c->stream));
  CK(cudaStreamWaitEvent(c->copyStream, c->evSnap, 0));
  if (x) CK(cudaMemcpyAsync(x, sx, 3 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->copyStream));
  if (R) CK(cudaMemcpyAsync(R, sR, 9 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->copyStream));
  if (v) CK(cudaMemcpyAsync(v, sv, 3 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->copyStream));
  if (omega) CK(cudaMemcpyAsync(omega, sw, 3 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->copyStream));
  if (sleeping) CK(cudaMemcpyAsync(sleeping, ssl, nb * sizeof(int), cudaMemcpyDeviceToHost, c->copyStream));
  if (collection) CK(cudaMemcpyAsync(collection, sco, nb * sizeof(int), cudaMemcpyDeviceToHost, c->copyStream));
  CK(cudaEventRecord(c->evCopied, c->copyStream));
  c->copyPending = true;
}
int am3d_download_bodies_async(am3d_ctx* c, double* x, double* R, double* v, double* omega, int32_t* sleeping, int32_t* collection) {
  API_BEGIN(c)
  downloadBodies(c, x, R, v, omega, sleeping, collection);
  API_END(c)
}
int am3d_wait_download(am3d_ctx* c) {
  API_BEGIN(c)
  if (c->copyPending) { CK(cudaEventSynchronize(c->evCopied)); c->copyPending = false; }
  API_END(c)
}
int am3d_download_bodies(am3d_ctx* c, double* x, double* R, double* v, double* omega, int32_t* sleeping, int32_t* collection) {
  API_BEGIN(c)
  downloadBodies(c, x, R, v, omega, sleeping, collection);
  CK(cudaEventSynchronize(c->evCopied));
  c->copyPending = false;
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
  // world-frame inertia follows the pose (RigidBody.updateRotationalInertiaFromTransformation :311-321)
  LAUNCH(c, k_update_inertia, nblk(nb), BLK, nb, c->flags.p, c->R.p, c->jinv0.p, c->mA0.p, c->jinv.p, c->mA.p);
  CK(cudaStreamSynchronize(c->stream));
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

int am3d_set_body_sleeping(am3d_ctx* c, int body, int sleeping) {
  API_BEGIN(c)
  if (!c->haveScene || body < 0 || body >= c->NB) throw AmError(AM3D_EINVAL, "bad body index");
  int fl;
  CK(cudaMemcpy(&fl, c->flags.p + body, sizeof(int), cudaMemcpyDeviceToHost));
  fl = sleeping ? (fl | AM3D_F_SLEEPING) : (fl & ~AM3D_F_SLEEPING);
  CK(cudaMemcpy(c->flags.p + body, &fl, sizeof(int), cudaMemcpyHostToDevice));
  API_END(c)
}

// LCPApp3D.java:936-947 (key 7): `if (b.magnetic) b.activateMagnet = !b.activateMagnet` for every body and every member of a
// collection; PGS.java:119,150,167 then leaves the multipliers of a contact with an active magnet unclamped.  The flag is
// read when the solve order is set up (k_group_setup), i.e. from the next step on.  A body that is not magnetic keeps its flags.
int am3d_set_body_magnet(am3d_ctx* c, int body, int active) {
  API_BEGIN(c)
  if (!c->haveScene || body < 0 || body >= c->NB) throw AmError(AM3D_EINVAL, "bad body index");
  int fl;
  CK(cudaMemcpy(&fl, c->flags.p + body, sizeof(int), cudaMemcpyDeviceToHost));
  if (fl & AM3D_F_MAGNETIC) {
    fl = active ? (fl | AM3D_F_MAGNET_ACTIVE) : (fl & ~AM3D_F_MAGNET_ACTIVE);
    CK(cudaMemcpy(c->flags.p + body, &fl, sizeof(int), cudaMemcpyHostToDevice));
  }
  API_END(c)
}

// one body: state, world-frame inertia, flags as in the scene blob minus DORMANT, list position at the end
__global__ void k_activate_body(int b, int flagsOn, long long stampNew, const double* __restrict__ st /* x3 R9 v3 w3 */, double* __restrict__ x,
                                double* __restrict__ R, double* __restrict__ v, double* __restrict__ w, int* __restrict__ flags,
                                long long* __restrict__ stamp, int* __restrict__ metricCount, const double* __restrict__ jinv0,
                                const double* __restrict__ mA0, double* __restrict__ jinv, double* __restrict__ mA) {
  if (blockIdx.x || threadIdx.x) return;
  for (int k = 0; k < 3; k++) { x[3 * b + k] = st[k]; v[3 * b + k] = st[12 + k]; w[3 * b + k] = st[15 + k]; }
  for (int k = 0; k < 9; k++) R[9 * b + k] = st[3 + k];
  flags[b] = flagsOn;
  stamp[b] = stampNew;
  metricCount[b] = 0;
  if (!(flagsOn & AM3D_F_PINNED)) {  // RigidBody.updateRotationalInertiaFromTransformation :311-321
    m3 Rm = ldm(R + 9 * b);
    stm(mA + 9 * b, rm0rt(Rm, ldm(mA0 + 9 * b)));
    stm(jinv + 9 * b, rm0rt(Rm, ldm(jinv0 + 9 * b)));
  }
}
int am3d_activate_body(am3d_ctx* c, int body, const double x[3], const double R[9], const double v[3], const double omega[3]) {
  API_BEGIN(c)
  if (!c->haveScene || body < 0 || body >= c->NB || !x) throw AmError(AM3D_EINVAL, "bad body index / null position");
  int fl;
  CK(cudaMemcpy(&fl, c->flags.p + body, sizeof(int), cudaMemcpyDeviceToHost));
  if (!(fl & AM3D_F_DORMANT)) throw AmError(AM3D_ESTATE, "body is already part of the simulation");
  double st[18] = {x[0], x[1], x[2], 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0};
  if (R) for (int k = 0; k < 9; k++) st[3 + k] = R[k];
  if (v) for (int k = 0; k < 3; k++) st[12 + k] = v[k];
  if (omega) for (int k = 0; k < 3; k++) st[15 + k] = omega[k];
  c->pokeV.ensure(18);
  CK(cudaMemcpyAsync(c->pokeV.p, st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
  int on = c->H.body_flags[body] & ~(AM3D_F_DORMANT | AM3D_F_SLEEPING);
  if (c->H.body_type[body] == AM3D_BODY_PLANE) on |= AM3D_F_PINNED;
  LAUNCH(c, k_activate_body, 1, 32, body, on, c->nextStamp, c->pokeV.p, c->x.p, c->R.p, c->v.p, c->w.p, c->flags.p, c->stamp.p, c->metricCount.p,
         c->jinv0.p, c->mA0.p, c->jinv.p, c->mA.p);
  c->nextStamp++;
  c->nDormant--;
  CK(cudaStreamSynchronize(c->stream));
  API_END(c)
}
int am3d_remove_body(am3d_ctx* c, int body) {
  API_BEGIN(c)
  if (!c->haveScene || body < 0 || body >= c->NB) throw AmError(AM3D_EINVAL, "bad body index");
  int fl, par;
  CK(cudaMemcpy(&fl, c->flags.p + body, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&par, c->parent.p + body, sizeof(int), cudaMemcpyDeviceToHost));
  if (fl & AM3D_F_DORMANT) throw AmError(AM3D_ESTATE, "body is not part of the simulation");
  if (par >= 0) throw AmError(AM3D_ESTATE, "body is part of a collection");
  fl |= AM3D_F_DORMANT | AM3D_F_PINNED | AM3D_F_SLEEPING;
  CK(cudaMemcpy(c->flags.p + body, &fl, sizeof(int), cudaMemcpyHostToDevice));
  c->nDormant++;
  API_END(c)
}
int am3d_set_mouse_spring(am3d_ctx* c, int body, const double grabPointB[3], const double pointW[3], double stiffness, double damping,
                          int apply_at_com) {
  API_BEGIN(c)
  if (!c->haveScene || body >= c->NB) throw AmError(AM3D_EINVAL, "bad body index");
  if (body >= 0 && (!grabPointB || !pointW)) throw AmError(AM3D_EINVAL, "null points");
  MouseState ms;
  CK(cudaMemcpy(&ms, c->mouse.p, sizeof(ms), cudaMemcpyDeviceToHost));
  ms.springBody = body < 0 ? -1 : body;
  if (body >= 0) {
    for (int k = 0; k < 3; k++) { ms.grabB[k] = grabPointB[k]; ms.pointW[k] = pointW[k]; }
    ms.k = stiffness; ms.c = damping; ms.atCOM = apply_at_com;
    c->mouseUsed = true;
  }
  CK(cudaMemcpy(c->mouse.p, &ms, sizeof(ms), cudaMemcpyHostToDevice));
  API_END(c)
}
int am3d_apply_impulse(am3d_ctx* c, int body, const double pickedPointB[3], const double endPointW[3], double scale) {
  API_BEGIN(c)
  if (!c->haveScene || body < 0 || body >= c->NB || !pickedPointB || !endPointW) throw AmError(AM3D_EINVAL, "bad body index / null points");
  MouseState ms;
  CK(cudaMemcpy(&ms, c->mouse.p, sizeof(ms), cudaMemcpyDeviceToHost));
  ms.impBody = body; ms.impPhase = 1; ms.impScale = scale;
  for (int k = 0; k < 3; k++) { ms.impPointB[k] = pickedPointB[k]; ms.impEndW[k] = endPointW[k]; }
  CK(cudaMemcpy(c->mouse.p, &ms, sizeof(ms), cudaMemcpyHostToDevice));
  c->mouseUsed = true;
  API_END(c)
}

// The per-body pokes travel on an upload stream and are added to the velocities by the NEXT step after its contact
// detection - detection reads positions only, so the upload (48 B per body) runs beside it instead of in front of the step -
// or by whatever entry point is called first (every other call flushes a pending poke, so velocities read back or written
// in between see it applied).  The host buffers must stay untouched until that step (or call) has returned, as with any
// cudaMemcpyAsync from pinned memory.
int am3d_add_velocities(am3d_ctx* c, const double* dv, const double* domega) {
  API_BEGIN(c)   // (a poke still pending from an earlier call is applied first)
  if (!c->haveScene || !dv || !domega) throw AmError(AM3D_EINVAL, "no scene / null buffers");
  int nb = c->NB;
  ensureCopyStream(c);
  c->pokeV.ensure(3 * (size_t)nb); c->pokeW.ensure(3 * (size_t)nb);
  CK(cudaEventRecord(c->evMain, c->stream));             // the staging buffers exist / were read by the last poke
  CK(cudaStreamWaitEvent(c->upStream, c->evMain, 0));
  CK(cudaMemcpyAsync(c->pokeV.p, dv, 3 * (size_t)nb * sizeof(double), cudaMemcpyHostToDevice, c->upStream));
  CK(cudaMemcpyAsync(c->pokeW.p, domega, 3 * (size_t)nb * sizeof(double), cudaMemcpyHostToDevice, c->upStream));
  CK(cudaEventRecord(c->evPoke, c->upStream));
  c->pokesPending = true;
  API_END(c)
}

int am3d_num_contacts(am3d_ctx* c, int include_internal) {
  if (!c) return AM3D_EINVAL;
  waitPending(c);
  return c->cur.n + (include_internal ? c->icon.n : 0);
}

int am3d_download_contacts(am3d_ctx* c, am3d_contact* out, int capacity, int include_internal, int* count) {
  API_BEGIN(c)
  int n = std::min(c->cur.n, capacity);
  fetchContacts(c, c->cur, n, out, 0);
  int m = 0;
  if (include_internal) {
    m = std::min(c->icon.n, capacity - n);
    fetchContacts(c, c->icon, m, out + n, 1);
  }
  if (count) *count = n + m;
  API_END(c)
}

// merge / unmerge decisions so far: (step, kind 0 = pair became internal / 1 = pair left a collection, bodyLo, bodyHi)
int am3d_num_events(am3d_ctx* c) { if (c) waitPending(c); return c ? (int)(c->events.size() / 4) : AM3D_EINVAL; }
int am3d_download_events(am3d_ctx* c, int32_t* out, int capacity, int* count) {
  API_BEGIN(c)
  int n = std::min((int)(c->events.size() / 4), capacity);
  if (count) *count = n;
  if (n) memcpy(out, c->events.data(), (size_t)n * 4 * sizeof(int32_t));
  API_END(c)
}

// tests: keep the Gauss-Seidel sequence of every solve so that the CPU oracle can replay it
int am3d_record_orders(am3d_ctx* c, int on) {
  if (!c) return AM3D_EINVAL;
  c->recordOrders = on != 0;
  return AM3D_OK;
}
// which = 0: last full solve, 1: last single sweep.  Only the identity fields of `out` are filled.
int am3d_download_order(am3d_ctx* c, int which, am3d_contact* out, int capacity, int* count) {
  API_BEGIN(c)
  std::vector<am3d_contact>& keys = which == 1 ? c->orderSweepKeys : which == 2 ? c->orderPostKeys : c->orderFullKeys;
  int n = (int)keys.size();
  if (count) *count = n;
  if (!out) return AM3D_OK;  // size query
  if (n > capacity) throw AmError(AM3D_EINVAL, "capacity too small");
  if (n) memcpy(out, keys.data(), (size_t)n * sizeof(am3d_contact));
  API_END(c)
}

int am3d_num_bpcs(am3d_ctx* c) { if (c) waitPending(c); return c ? c->bpPrev.n : AM3D_EINVAL; }
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

int am3d_num_internal_bpcs(am3d_ctx* c) { if (c) waitPending(c); return c ? c->ibp.n : AM3D_EINVAL; }
int am3d_download_internal_bpcs(am3d_ctx* c, am3d_bpc* out, int capacity, int* count) {
  API_BEGIN(c)
  int n = std::min(c->ibp.n, capacity);
  if (count) *count = n;
  if (n > 0) {
    std::vector<int> b1(n), b2(n), nm(n), cnt(n), al(n);
    CK(cudaMemcpyAsync(b1.data(), c->ibp.b1.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(b2.data(), c->ibp.b2.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(nm.data(), c->ibp.nMetric.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(cnt.data(), c->ibp.count.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(al.data(), c->ibp.alive.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) {
      memset(&out[i], 0, sizeof(am3d_bpc));
      out[i].body1 = b1[i]; out[i].body2 = b2[i]; out[i].in_collection = al[i]; out[i].n_contacts = cnt[i]; out[i].n_metric = nm[i];
    }
  }
  API_END(c)
}

int am3d_get_timings(am3d_ctx* c, am3d_timings* t) {
  if (!c || !t) return AM3D_EINVAL;
  waitPending(c);
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
  applyExternalForces(c, dt);  // clear + gravity + springs, deltaV = 0
  runSolve(c, dt, false);
  c->orderFullKeys.clear();
  if (c->recordOrders) snapshotOrder(c, 0);
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
  CK(cudaMemcpy2DAsync(dv, 6 * sizeof(double), c->dv.p, DVS * sizeof(double), 6 * sizeof(double), c->NB, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END(c)
}

int am3d_download_solve_order(am3d_ctx* c, int32_t* order, int capacity, int* count) {
  API_BEGIN(c)
  int n = c->cur.n;
  if (count) *count = n;
  if (n > capacity) throw AmError(AM3D_EINVAL, "capacity too small");
  if (n > 0) {
    if (c->lastSolveSweep || c->nPairsSolve != c->bp.n || c->nGroups == 0) throw AmError(AM3D_ESTATE, "no full solve has been run on the current contacts");
    // order[k] = canonical contact index solved k-th
    CK(cudaMemcpyAsync(order, c->scSrc.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  API_END(c)
}

int am3d_stats(am3d_ctx* c, double* out /* [4]: kernel launches, solve launches, row updates, solve kernel seconds */) {
  if (!c || !out) return AM3D_EINVAL;
  waitPending(c);
  out[0] = (double)c->kernelLaunches; out[1] = (double)c->solveLaunches; out[2] = c->rowUpdates; out[3] = c->solveSeconds;
  return AM3D_OK;
}

int am3d_set_option(am3d_ctx* c, const char* name, double value) {
  API_BEGIN(c)
  if (!name) throw AmError(AM3D_EINVAL, "null option name");
  if (!strcmp(name, "hub_min_degree")) c->hubMin = (int)value;
  else if (!strcmp(name, "pgs_persistent")) c->usePersistent = (int)value;
  else if (!strcmp(name, "record_events")) c->recordEvents = value != 0;
  else if (!strcmp(name, "giant_warps")) c->useGiantWarps = (int)value;
  else if (!strcmp(name, "pgs_fast_rows")) c->fastRows = value != 0;
  else if (!strcmp(name, "tree_split")) c->treeSplit = value != 0;
  else if (!strcmp(name, "scene_bfs")) c->useSceneBfs = value != 0;
  else if (!strcmp(name, "own_primitives")) c->ownPrimitives = value != 0;
  else if (!strcmp(name, "pgs_tail_fusion")) c->useTailFusion = value != 0;
  else if (!strcmp(name, "pgs_clusters")) c->useClusters = (int)value;
  else if (!strcmp(name, "giant_chunk")) c->giantChunk = (int)value;
  else if (!strcmp(name, "merge_exact_max_pairs")) c->mergeExactMax = (int)value;
  else throw AmError(AM3D_EINVAL, std::string("unknown option ") + name);
  API_END(c)
}

// device-side timer on the context's own stream: am3d_mark(slot 0/1) + am3d_elapsed_ms
int am3d_mark(am3d_ctx* c, int slot) {
  API_BEGIN(c)
  if (slot < 0 || slot > 1) throw AmError(AM3D_EINVAL, "slot must be 0 or 1");
  CK(cudaEventRecord(c->ev[12 + slot], c->stream));
  API_END(c)
}
int am3d_elapsed_ms(am3d_ctx* c, double* ms) {
  API_BEGIN(c)
  CK(cudaEventSynchronize(c->ev[13]));
  float f = 0;
  CK(cudaEventElapsedTime(&f, c->ev[12], c->ev[13]));
  *ms = f;
  API_END(c)
}

// debug / drawing: state of one collection slot: x[3] R[9] v[3] w[3] mass minv jinv[9] mA[9] flags alive count stamp
int am3d_download_collection(am3d_ctx* c, int slot, double* out /* [42] */) {
  API_BEGIN(c)
  int nc = c->NS - c->NB;
  if (slot < 0 || slot >= nc) throw AmError(AM3D_EINVAL, "bad collection slot");
  int s = c->NB + slot;
  auto g = [&](double* d, const double* p, int n) { CK(cudaMemcpy(d, p, n * sizeof(double), cudaMemcpyDeviceToHost)); };
  CK(cudaStreamSynchronize(c->stream));
  g(out, c->x.p + 3 * s, 3); g(out + 3, c->R.p + 9 * s, 9); g(out + 12, c->v.p + 3 * s, 3); g(out + 15, c->w.p + 3 * s, 3);
  g(out + 18, c->mass.p + s, 1); g(out + 19, c->minv.p + s, 1); g(out + 20, c->jinv.p + 9 * s, 9); g(out + 29, c->mA.p + 9 * s, 9);
  int fl, al, cnt;
  long long st;
  CK(cudaMemcpy(&fl, c->flags.p + s, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&al, c->collAlive.p + slot, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cnt, c->collCount.p + slot, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, c->stamp.p + s, sizeof(long long), cudaMemcpyDeviceToHost));
  out[38] = fl; out[39] = al; out[40] = cnt; out[41] = (double)st;
  API_END(c)
}

__global__ void k_list_order(int nb, const int* __restrict__ parent, const long long* __restrict__ stamp, long long* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  int p = parent[i];
  out[i] = stamp[p >= 0 ? p : i];
}
int am3d_download_list_order(am3d_ctx* c, int64_t* out) {
  API_BEGIN(c)
  if (!c->haveScene || !out) throw AmError(AM3D_EINVAL, "no scene / null buffer");
  DevBuf<long long> tmp;
  tmp.ensure(c->NB + 1);
  LAUNCH(c, k_list_order, nblk(c->NB), BLK, c->NB, c->parent.p, c->stamp.p, tmp.p);
  CK(cudaMemcpyAsync(out, tmp.p, c->NB * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END(c)
}

// debug: assembled solver row of one external contact after the last full solve:
// out = dirs[9] r1r2[6] b[3] D[3] lam[3] mass[20] sgB1 sgB2 start count color pos
int am3d_debug_solve_row(am3d_ctx* c, int contact, double* out /* [50] */) {
  API_BEGIN(c)
  int n = c->lastSolveN;
  std::vector<int> src(n);
  CK(cudaMemcpy(src.data(), c->scSrc.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  int idx = -1;
  for (int k = 0; k < n; k++) if (src[k] == contact) idx = k;
  if (idx < 0) throw AmError(AM3D_EINVAL, "contact not in the last solve");
  if (c->nGroups != c->nPairsSolve) throw AmError(AM3D_ESTATE, "the last solve cut body pairs into chunks: no per-pair row to show");
  auto g = [&](double* d, const double* p, int m) { CK(cudaMemcpy(d, p, m * sizeof(double), cudaMemcpyDeviceToHost)); };
  g(out, c->scP.p + 24 * (size_t)idx, 24);
  int bpc;
  CK(cudaMemcpy(&bpc, c->cur.bpc.p + contact, sizeof(int), cudaMemcpyDeviceToHost));
  int pos, col, b1, b2, st, cnt;
  CK(cudaMemcpy(&pos, c->grpPos.p + bpc, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&col, c->grpColor.p + bpc, sizeof(int), cudaMemcpyDeviceToHost));
  g(out + 24, c->sgMass.p + 20 * pos, 20);
  CK(cudaMemcpy(&b1, c->sgB1.p + pos, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&b2, c->sgB2.p + pos, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, c->sgStart.p + pos, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cnt, c->sgCount.p + pos, sizeof(int), cudaMemcpyDeviceToHost));
  out[44] = b1; out[45] = b2; out[46] = st; out[47] = cnt; out[48] = col; out[49] = pos;
  API_END(c)
}

}  // extern "C"
