// Host orchestration of the two PGS solves (full solve and single sweep) and of merge / unmerge.
#pragma once
#include "am3d_host_util.cuh"
#include "am3d_sort.cuh"
#include "am3d_merge.cuh"
#include "am3d_step.cuh"

static SolveArrays solveArrays(am3d_ctx* c) {
  return SolveArrays{c->sgB1.p, c->sgB2.p, c->sgStart.p, c->sgCount.p, c->sgFlags.p, c->sgBpc.p, c->sgScene.p, c->sceneState.p, c->sgMass.p, c->sgMu.p,
                     c->scP.p, c->scSrc.p, c->scState.p, c->hubDelta.p};
}
static ContactPtrs contactPtrs(ContactSet& S) {
  return ContactPtrs{S.b1.p, S.b2.p, S.s1.p, S.s2.p, S.bv1.p, S.bv2.p, S.info.p, S.leaf.p, S.bpc.p, S.state.p, S.isNew.p,
                     S.key0.p, S.key1.p, S.pW.p, S.nW.p, S.t1W.p, S.t2W.p, S.pB1.p, S.nB1.p, S.t1B1.p, S.t2B1.p, S.viol.p,
                     S.prevViol.p, S.lam.p, S.lamWarm.p};
}

// colour the groups (body pairs) so that no two groups of a colour share a non-pinned solver body
// and order them by (layer, colour) phases; layer = nullptr for the full solve
static void colourGroups(am3d_ctx* c, int ng, const int* gb1, const int* gb2, const int* gcount, const int* lead, int inCollection,
                         const int* layer, int layerBits, bool giantsPossible) {
  c->grpColor.ensure(ng + 1); c->grpPrio.ensure(ng + 1); c->grpSb1.ensure(ng + 1); c->grpSb2.ensure(ng + 1);
  c->grpPos.ensure(ng + 1); c->grpOrder.ensure(ng + 1); c->grpKey.ensure(ng + 1); c->grpKeySorted.ensure(ng + 1); c->grpVal.ensure(ng + 1);
  c->grpDegree.ensure(c->NS + 1); c->grpHubMask.ensure(ng + 1);
  CK(cudaMemsetAsync(c->grpDegree.p, 0, c->NS * sizeof(int), c->stream));
  CK(cudaMemsetAsync(c->counters.p + 6, 0, sizeof(int), c->stream));
  LAUNCH(c, k_grp_init, nblk(ng), BLK, ng, gb1, gb2, c->parent.p, c->flags.p, gcount, c->bodyLocal.p, lead, inCollection, c->grpSb1.p, c->grpSb2.p,
         c->grpPrio.p, c->grpColor.p, c->grpDegree.p);
  LAUNCH(c, k_grp_hubs, nblk(ng), BLK, ng, c->grpSb1.p, c->grpSb2.p, c->grpDegree.p, c->hubMin, c->grpHubMask.p, c->counters.p + 6);
  CK(cudaMemsetAsync(c->bodyBest.p, 0, c->NS * sizeof(unsigned long long), c->stream));
  CK(cudaMemsetAsync(c->bodyMask.p, 0, c->NS * sizeof(unsigned long long), c->stream));
  CK(cudaMemsetAsync(c->counters.p + 3, 0, sizeof(int), c->stream));
  int page = 0;
  const int maxPages = 64;
  if (c->colorBlocks <= 0) throw AmError(AM3D_ECUDA, "cooperative launch unsupported");
  {  // the rounds run on the device (one cooperative launch, no read-back per round)
    CK(cudaMemsetAsync(c->colorCtl.p, 0, 8 * sizeof(int), c->stream));
    int ngv = ng, nsv = c->NS, mp = maxPages;
    const int *s1 = c->grpSb1.p, *s2 = c->grpSb2.p, *hm = c->grpHubMask.p;
    const unsigned long long* pr = c->grpPrio.p;
    int* col = c->grpColor.p;
    unsigned long long *bb = c->bodyBest.p, *bm = c->bodyMask.p;
    int* ctl = c->colorCtl.p;
    void* args[] = {&ngv, &nsv, &mp, &s1, &s2, &hm, &pr, &col, &bb, &bm, &ctl};
    int blocks = std::min(c->colorBlocks, std::max(1, nblk(ng, 256)));
    CK(cudaLaunchCooperativeKernel((const void*)k_color_coop, dim3(blocks), dim3(256), args, 0, c->stream));
    c->kernelLaunches++;
    int out[2];
    readBack(c, out, c->colorCtl.p + 4, 2);
    if (out[1]) throw AmError(AM3D_ECAPACITY, "more than 4096 colours needed");
    page = out[0] - 1;
  }
  int maxColors = (page + 1) * 64;
  int endBit = 20 + (layer ? layerBits : 0);
  if (!layer) endBit = 8 + bitsFor((unsigned long long)maxColors);
  // many small scenes: partitions of consecutive scenes, one thread-block cluster each (k_pgs_cluster)
  int nScenes = c->H.nscenes;
  c->nPart = 0;
  int partShift = 0;
  if (c->useClusters && c->maxClusters > 0 && nScenes >= 2 * c->maxClusters && nScenes <= 64 * c->maxClusters && !giantsPossible) {
    c->nPart = c->maxClusters;
    partShift = endBit;
    endBit += bitsFor((unsigned long long)c->nPart);
  }
  LAUNCH(c, k_color_sortkey, nblk(ng), BLK, ng, c->grpColor.p, gcount, layer, gb1, c->scene.p, partShift, c->nPart, nScenes, c->grpKey.p,
         c->grpVal.p);
  sortPairs(c, c->grpKey.p, c->grpKeySorted.p, c->grpVal.p, c->grpOrder.p, ng, 0, endBit);
  // phases = runs of equal (layer, colour) in the sorted list
  c->phaseHead.ensure(ng + 2); c->phaseScan.ensure(ng + 2); c->sgPhase.ensure(ng + 2);
  LAUNCH(c, k_phase_heads, nblk(ng), BLK, ng, c->grpKeySorted.p, c->phaseHead.p);
  int nPhases = scanTotal(c, c->phaseHead, c->phaseScan, ng);
  c->dColorStart.ensure(nPhases + 2);
  LAUNCH(c, k_phase_fill, nblk(ng), BLK, ng, c->phaseHead.p, c->phaseScan.p, c->dColorStart.p, c->sgPhase.p);
  c->colorStart.resize(nPhases + 1);
  readBack(c, c->colorStart.data(), c->dColorStart.p, (size_t)nPhases + 1);
  c->nColors = nPhases;
  c->nGroups = ng;
  if (c->nPart > 0) {
    c->partRange.ensure(2 * (size_t)c->nPart + 2);
    CK(cudaMemsetAsync(c->partRange.p, 0, 2 * (size_t)c->nPart * sizeof(int), c->stream));
    LAUNCH(c, k_part_phases, nblk(nPhases), BLK, nPhases, c->dColorStart.p, c->grpKeySorted.p, partShift, c->partRange.p);
    if ((int)c->hPartSceneStart.size() != c->nPart + 1 || c->hPartScenes != nScenes) {  // scenes of partition p: those with s * nPart / nScenes == p
      c->hPartSceneStart.assign(c->nPart + 1, nScenes);
      for (int s = nScenes - 1; s >= 0; s--) c->hPartSceneStart[(long long)s * c->nPart / nScenes] = s;
      for (int k = c->nPart - 1; k >= 0; k--) c->hPartSceneStart[k] = std::min(c->hPartSceneStart[k], c->hPartSceneStart[k + 1]);
      c->hPartScenes = nScenes;
      c->hPartCount.resize(c->nPart);
      for (int k = 0; k < c->nPart; k++) c->hPartCount[k] = c->hPartSceneStart[k + 1] - c->hPartSceneStart[k];
      c->partSceneStart.ensure(c->nPart + 2);
      CK(cudaMemcpyAsync(c->partSceneStart.p, c->hPartSceneStart.data(), (c->nPart + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
  }
}

// One contact set taking part in a solve: its groups start at groupOffset in the group list.
struct SolveSet {
  ContactSet* S;
  int groupOffset;
  int setId;
  bool writeLambda;
};

// PGS.solve (PGS.java:73-194).  sweep = false: the full solve over the external contacts with collections as
// solver bodies (CollisionProcessor.solveLCP :108-137); sweep = true: the single sweep over external + internal
// contacts with every body on its own (updateInCollections :232-303).
// post = true: the position-level solve of postStabilization (RigidBodySystem.java:354-377, PGS.java:86-89)
static void runSolve(am3d_ctx* c, double dt, bool sweep, bool post = false) {
  const am3d_params& P = c->P;
  int nExt = c->bp.n, nInt = sweep ? c->ibp.n : 0;
  int ng = nExt + nInt;
  int ncExt = c->cur.n, ncInt = sweep ? c->icon.n : 0;
  int nc = ncExt + ncInt;
  if (!sweep && !post) { c->T.pgs_iterations = 0; c->T.pgs_colors = 0; c->T.pgs_kernel_time = 0; }
  c->lastSolveSweep = sweep;
  c->lastSolveN = 0;
  (sweep ? c->orderSweep : post ? c->orderPost : c->orderFull).clear();
  if (nc == 0 || ng == 0) return;
  const int *gb1, *gb2, *gcount, *gstart;
  c->swB1.ensure(ng + 1); c->swB2.ensure(ng + 1); c->swCount.ensure(ng + 1); c->swStart.ensure(ng + 1);
  c->swAsleep.ensure(ng + 1);
  LAUNCH(c, k_sweep_groups, nblk(ng), BLK, nExt, nInt, c->cur.b1.p, c->cur.b2.p, c->bp.count.p, c->bp.start.p, c->ibp.b1.p,
         c->icon.b1.p, c->icon.b2.p, c->ibp.count.p, c->ibp.start.p, c->ibp.alive.p, c->parent.p, c->flags.p, c->swB1.p, c->swB2.p,
         c->swCount.p, c->swStart.p, c->swAsleep.p);
  gb1 = c->swB1.p; gb2 = c->swB2.p; gcount = c->swCount.p; gstart = c->swStart.p;
  const int* layer = nullptr;
  int layerBits = 0;
  if (sweep && c->P.organize_contacts) {
    // getOrganizedContacts (CollisionProcessor.java:346-441): breadth-first layers from the pairs with new contacts
    CK(cudaEventRecord(c->ev[20], c->stream));
    c->grpLayer.ensure(ng + 1); c->bodyLevel.ensure(c->NB + 1); c->bfsRound.ensure(8);
    CK(cudaMemsetAsync(c->grpLayer.p, 0x7f, ng * sizeof(int), c->stream));
    CK(cudaMemsetAsync(c->bodyLevel.p, 0x7f, c->NB * sizeof(int), c->stream));
    CK(cudaMemsetAsync(c->bfsRound.p, 0, 8 * sizeof(int), c->stream));
    if (ncExt > 0) LAUNCH(c, k_bfs_seed, nblk(ncExt), BLK, ncExt, c->cur.isNew.p, c->cur.bpc.p, c->grpLayer.p);
    if (c->mouseUsed) {
      LAUNCH(c, k_bfs_seed_picked, nblk(ng), BLK, ng, gb1, gb2, gcount, c->picked.p, c->grpLayer.p);
      CK(cudaMemsetAsync(c->picked.p, 0, c->NB * sizeof(int), c->stream));  // b.picked = false (:371, :380)
    }
    // batched scenes, up to the size from which a pass over all groups per layer is cheaper than bucketing them (measured:
    // 512 scenes 10.97 -> 10.77 ms per step on average, 4096 scenes 49.58 -> 49.73): a CTA per scene, __syncthreads() between the layers
    if (c->useSceneBfs && c->H.nscenes >= 32 && c->H.nscenes <= 64 * std::max(1, c->maxClusters)) {
      int nsc = c->H.nscenes;
      c->sceneCnt.ensure(nsc + 2); c->sceneStart.ensure(nsc + 2); c->sceneCursor.ensure(nsc + 2); c->sceneList.ensure(ng + 1);
      CK(cudaMemsetAsync(c->sceneCnt.p, 0, (nsc + 1) * sizeof(int), c->stream));
      CK(cudaMemsetAsync(c->sceneCursor.p, 0, (nsc + 1) * sizeof(int), c->stream));
      LAUNCH(c, k_scene_count, nblk(ng), BLK, ng, gb1, c->scene.p, c->sceneCnt.p);
      exclusiveSum(c, c->sceneCnt.p, c->sceneStart.p, nsc + 1);
      LAUNCH(c, k_scene_fill, nblk(ng), BLK, ng, gb1, c->scene.p, c->sceneStart.p, c->sceneCursor.p, c->sceneList.p);
      LAUNCH(c, k_bfs_scenes, nsc, 256, c->sceneStart.p, c->sceneList.p, gb1, gb2, gcount, c->grpLayer.p, c->bodyLevel.p, c->bfsRound.p);
    } else {
      int ngv = ng;
      int *gl = c->grpLayer.p, *bl = c->bodyLevel.p, *rd = c->bfsRound.p;
      void* args[] = {&ngv, &gb1, &gb2, &gcount, &gl, &bl, &rd};
      int blocks = std::min(c->bfsBlocks, std::max(1, nblk(ng, 256)));
      if (c->bfsBlocks <= 0) throw AmError(AM3D_ECUDA, "cooperative launch unsupported");
      CK(cudaLaunchCooperativeKernel((const void*)k_bfs_layers, dim3(blocks), dim3(256), args, 0, c->stream));
      c->kernelLaunches++;
    }
    LAUNCH(c, k_bfs_finalize, nblk(ng), BLK, ng, nExt, c->swAsleep.p, c->bfsRound.p, c->grpLayer.p, c->swCount.p);
    CK(cudaEventRecord(c->ev[21], c->stream));
    c->orderingTimed = true;
    int deepest = readInt(c, c->bfsRound.p + 3);
    layer = c->grpLayer.p;
    layerBits = bitsFor((unsigned long long)deepest + 2);
    if (layerBits > 44) throw AmError(AM3D_ECAPACITY, "too many sweep layers");
  } else if (sweep) {
    // organize_contacts = false (CollisionProcessor.java:249-258): external contacts, then the internal contacts of
    // the awake collections
    c->grpLayer.ensure(ng + 1);
    LAUNCH(c, k_plain_layers, nblk(ng), BLK, ng, nExt, c->swAsleep.p, c->grpLayer.p, c->swCount.p);
    layer = c->grpLayer.p;
    layerBits = 1;
  }
  // pairs with hundreds of contacts exist only between sphere trees with many leaves: ask the last detection
  bool giantsPossible = false;
  if (c->NN > 1) giantsPossible = readInt(c, c->counters.p + 8 + c->bpSlot) > 64 || (sweep && c->icon.n > 0);
  // cut them into chunks
  const int np = ng;
  const int *pcount = gcount, *pstart = gstart;
  const int *chunkFirst = nullptr, *lead = nullptr;
  int chunkLen = 0;
  if (giantsPossible && c->giantChunk > 0) {
    c->chN.ensure(np + 2); c->chFirst.ensure(np + 2);
    LAUNCH(c, k_chunk_count, nblk(np), BLK, np, pcount, c->giantChunk, c->chN.p);
    int total = scanTotal(c, c->chN, c->chFirst, np);
    if (total > np) {
      ng = total;
      chunkLen = c->giantChunk;
      c->cgB1.ensure(ng + 1); c->cgB2.ensure(ng + 1); c->cgCount.ensure(ng + 1); c->cgStart.ensure(ng + 1); c->cgLayer.ensure(ng + 1); c->cgLead.ensure(ng + 1);
      LAUNCH(c, k_chunk_expand, nblk(np), BLK, np, chunkLen, c->chFirst.p, gb1, gb2, pcount, pstart, layer, c->cgB1.p, c->cgB2.p, c->cgCount.p,
             c->cgStart.p, c->cgLayer.p, c->cgLead.p);
      gb1 = c->cgB1.p; gb2 = c->cgB2.p; gcount = c->cgCount.p; gstart = c->cgStart.p;
      if (layer) layer = c->cgLayer.p;
      chunkFirst = c->chFirst.p; lead = c->cgLead.p;
    }
  }
  c->nPairsSolve = np;
  colourGroups(c, ng, gb1, gb2, gcount, lead, sweep ? 1 : 0, layer, layerBits, giantsPossible);
  c->sgB1.ensure(ng + 1); c->sgB2.ensure(ng + 1); c->sgStart.ensure(ng + 2); c->sgCount.ensure(ng + 2); c->sgFlags.ensure(ng + 1);
  c->sgBpc.ensure(ng + 1); c->sgMass.ensure(20 * (size_t)ng + 20); c->sgMu.ensure(ng + 1); c->sgScene.ensure(ng + 1);
  int nScenes = c->H.nscenes;
  c->sceneState.ensure((2 + MV_SLOTS) * (size_t)nScenes + 1);
  c->scP.ensure(24 * (size_t)nc + 48); c->scSrc.ensure(nc + 1); c->scState.ensure(nc + 1); c->hubDelta.ensure(12 * (size_t)ng + 12);
  SolveArrays S = solveArrays(c);
  LAUNCH(c, k_group_setup, nblk(ng), BLK, ng, c->grpOrder.p, c->grpSb1.p, c->grpSb2.p, gb1, gb2, gcount, c->minv.p, c->jinv.p,
         c->fric.p, c->flags.p, c->grpHubMask.p, P.friction_override, P.friction, c->scene.p, S, c->grpPos.p);
  int nSolve = scanTotal(c, c->sgCount, c->sgStart, ng);  // contacts that take part (sleeping collections excluded)
  c->lastSolveN = nSolve;
  // hub runs: (colour, hub body) -> the groups of that colour touching the hub, ascending
  c->nHubRuns = c->nHubEntries = 0;
  c->colorRunStart.assign(c->nColors + 1, 0);
  if (readInt(c, c->counters.p + 6) > 0) {
    c->hubN.ensure(ng + 2); c->hubScan.ensure(ng + 2);
    LAUNCH(c, k_hub_sides, nblk(ng), BLK, ng, c->sgFlags.p, c->hubN.p);
    int ne = scanTotal(c, c->hubN, c->hubScan, ng);
    c->nHubEntries = ne;
    c->hubKey.ensure(ne + 2); c->hubKeySorted.ensure(ne + 2); c->hubSlot.ensure(ne + 2); c->hubSlotSorted.ensure(ne + 2);
    c->hubHead.ensure(ne + 2); c->hubRunStart.ensure(ne + 2); c->hubRunBody.ensure(ne + 2); c->hubRunColor.ensure(ne + 2);
    int bitsG = bitsFor((unsigned long long)ng), bitsB = bitsFor((unsigned long long)c->NS), bitsP = bitsFor((unsigned long long)c->nColors);
    if (bitsG + bitsB + bitsP > 64) throw AmError(AM3D_ECAPACITY, "hub run keys: groups x solver bodies x phases exceed 64 bits");
    LAUNCH(c, k_hub_entries, nblk(ng), BLK, ng, c->sgFlags.p, c->sgB1.p, c->sgB2.p, c->sgPhase.p, c->hubScan.p, bitsG, bitsB,
           c->hubKey.p, c->hubSlot.p);
    sortPairs(c, c->hubKey.p, c->hubKeySorted.p, c->hubSlot.p, c->hubSlotSorted.p, ne, 0, bitsG + bitsB + bitsP);
    LAUNCH(c, k_hub_run_heads, nblk(ne), BLK, ne, c->hubKeySorted.p, bitsG, c->hubHead.p);
    int nr = scanTotal(c, c->hubHead, c->hubScan, ne);
    c->nHubRuns = nr;
    LAUNCH(c, k_hub_run_fill, nblk(ne), BLK, ne, c->hubKeySorted.p, c->hubHead.p, c->hubScan.p, bitsG, bitsB, c->hubRunStart.p, c->hubRunBody.p,
           c->hubRunColor.p);
    CK(cudaMemcpyAsync(c->hubRunStart.p + nr, &c->nHubEntries, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    std::vector<int> rc(nr);
    readBack(c, rc.data(), c->hubRunColor.p, (size_t)nr);
    std::vector<int> perColor(c->nColors, 0);
    for (int r = 0; r < nr; r++) perColor[rc[r]]++;
    for (int k = 0; k < c->nColors; k++) c->colorRunStart[k + 1] = c->colorRunStart[k] + perColor[k];
    c->hubDelta.ensure(12 * (size_t)ng + 12);
    S = solveArrays(c);
  }
  HubRuns HR{c->hubRunStart.p, c->hubRunBody.p, c->hubSlotSorted.p};
  SolveSet sets[2] = {{&c->cur, 0, 0, !sweep}, {&c->icon, nExt, 1, true}};
  int nsets = sweep ? 2 : 1;
  for (int k = 0; k < nsets; k++) {
    ContactSet& CS = *sets[k].S;
    if (CS.n == 0) continue;
    LAUNCH(c, k_assemble, nblk(CS.n, 128), 128, CS.n, CS.bpc.p, sets[k].groupOffset, sets[k].setId, pstart, pcount, chunkFirst, chunkLen, gstart, c->grpPos.p,
           c->sgStart.p, CS.b1.p, CS.b2.p, c->parent.p, sweep ? 1 : 0, CS.pW.p, CS.nW.p, CS.t1W.p, CS.t2W.p, CS.pB1.p, CS.nB1.p,
           CS.t1B1.p, CS.t2B1.p, CS.viol.p, CS.lam.p, CS.state.p, c->x.p, c->R.p, c->v.p, c->w.p, c->force.p, c->torque.p,
           c->minv.p, c->jinv.p, c->rest.p, dt,
           // CollisionProcessor.java:119: the velocity solve drops the Baumgarte term when post-stabilisation is on
           (!sweep && !post && P.enable_post_stabilization) ? 0.0 : P.feedback_stiffness, post ? 1 : 0, P.restitution_override, P.restitution, S);
  }
  // giant groups (sphere-tree pairs with hundreds of contacts) lead their phases and are solved one warp each
  std::vector<int> giants(c->nColors, 0);
  int nGiants = 0;
  if (giantsPossible && c->useGiantWarps) {
    c->phaseGiants.ensure(c->nColors + 1);
    CK(cudaMemsetAsync(c->phaseGiants.p, 0, (c->nColors + 1) * sizeof(int), c->stream));
    LAUNCH(c, k_phase_giants, nblk(ng), BLK, ng, c->sgCount.p, c->sgPhase.p, c->phaseGiants.p);
    readBack(c, giants.data(), c->phaseGiants.p, (size_t)c->nColors);
    for (int g : giants) nGiants += g;
  }
  // iterState: [1] every scene done, [2] largest iteration count, [4] contact-iterations, [6] scenes still iterating
  unsigned long long is0[8] = {0, 0, 0, 0, 0, 0, (unsigned long long)nScenes, 0};
  CK(cudaMemcpyAsync(c->iterState.p, is0, sizeof(is0), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(c->sceneState.p, 0, (2 + MV_SLOTS) * (size_t)nScenes * sizeof(int), c->stream));
  PgsParams PP{sweep ? 1.0 : P.omega, P.enable_compliance ? P.compliance : 0.0, sweep ? 1e-5 : P.tolerance, P.sliding_threshold, nScenes, sweep ? 0 : 1, c->fastRows};
  int iterations = sweep ? P.iterations_in_collection : P.iterations;
  CK(cudaEventRecord(c->ev[sweep ? 8 : 10], c->stream));
  // many small colours (hubs, batched scenes): one cooperative launch with grid barriers; few large colours: one
  // launch per colour (no barrier cost, full occupancy per launch)
  long long avgGroups = ng / std::max(1, c->nColors);
  bool persistent = nGiants == 0 && c->coopBlocks > 0 && c->usePersistent != 0 && (c->usePersistent == 2 || avgGroups < 4 * (long long)c->coopBlocks * 128);
  if (!sweep && !post) { c->T.pgs_kernel = (c->nPart > 0 && nGiants == 0) ? 2 : persistent ? 1 : 0; c->T.pgs_giant_groups = nGiants; }
  if (c->nPart > 0 && nGiants == 0) {
    bool hubs = c->nHubRuns > 0;
    if (hubs) {
      c->dColorRunStart.ensure(c->colorRunStart.size() + 1);
      CK(cudaMemcpyAsync(c->dColorRunStart.p, c->colorRunStart.data(), c->colorRunStart.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    c->partRemaining.ensure(c->nPart + 1);
    CK(cudaMemcpyAsync(c->partRemaining.p, c->hPartCount.data(), c->nPart * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(c->nPart * PGS_CLUSTER);
    cfg.blockDim = dim3(128);
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = PGS_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int *pr = c->partRange.p, *pss = c->partSceneStart.p, *dcs = c->dColorStart.p, *dcr = hubs ? c->dColorRunStart.p : nullptr;
    int* prem = c->partRemaining.p;
    double* dvp = c->dv.p;
    unsigned long long* isp = c->iterState.p;
    if (hubs) CK(cudaLaunchKernelEx(&cfg, k_pgs_cluster<true>, pr, pss, prem, dcs, dcr, HR, S, dvp, PP, iterations, isp));
    else CK(cudaLaunchKernelEx(&cfg, k_pgs_cluster<false>, pr, pss, prem, dcs, dcr, HR, S, dvp, PP, iterations, isp));
    c->kernelLaunches++;
    if (!sweep) c->solveLaunches++;
  } else if (persistent) {
    int nColors = c->nColors;
    const int* dcs = c->dColorStart.p;
    const int* dcr = nullptr;
    if (c->nHubRuns > 0) {
      c->dColorRunStart.ensure(c->colorRunStart.size() + 1);
      CK(cudaMemcpyAsync(c->dColorRunStart.p, c->colorRunStart.data(), c->colorRunStart.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
      dcr = c->dColorRunStart.p;
    }
    double* dvp = c->dv.p;
    unsigned long long* isp = c->iterState.p;
    void* args[] = {&nColors, &dcs, &dcr, &HR, &S, &dvp, &PP, &iterations, &isp};
    bool hubs = c->nHubRuns > 0;
    const void* kfn = hubs ? (const void*)k_pgs_persistent<true> : (const void*)k_pgs_persistent<false>;
    CK(cudaLaunchCooperativeKernel(kfn, dim3(c->coopBlocksV[hubs ? 1 : 0]), dim3(128), args, 0, c->stream));
    c->kernelLaunches++;
    if (!sweep) c->solveLaunches++;
  } else {
    const bool hubs = c->nHubRuns > 0;
    const size_t gsm = (size_t)GIANT_WARPS * 32 * GIANT_ROW * sizeof(double);
    auto giantLaunch = [&](int mode, int g0, int n, int last) {
      int grid = (n + GIANT_WARPS - 1) / GIANT_WARPS;
      if (mode == 0) {
        if (hubs) k_pgs_giant<0, true><<<grid, 32 * GIANT_WARPS, gsm, c->stream>>>(g0, g0 + n, S, c->dv.p, PP, last, c->iterState.p);
        else k_pgs_giant<0, false><<<grid, 32 * GIANT_WARPS, gsm, c->stream>>>(g0, g0 + n, S, c->dv.p, PP, last, c->iterState.p);
      } else {
        if (hubs) k_pgs_giant<1, true><<<grid, 32 * GIANT_WARPS, gsm, c->stream>>>(g0, g0 + n, S, c->dv.p, PP, last, c->iterState.p);
        else k_pgs_giant<1, false><<<grid, 32 * GIANT_WARPS, gsm, c->stream>>>(g0, g0 + n, S, c->dv.p, PP, last, c->iterState.p);
      }
      CK(cudaGetLastError());
      c->kernelLaunches++;
    };
    // tail phases with at most one group per scene (batched scenes): one launch for all of them, a thread per scene
    int cTail = c->nColors;
    if (c->useTailFusion && !hubs && nScenes >= 8) {
      while (cTail > 0 && giants[cTail - 1] == 0 && c->colorStart[cTail] - c->colorStart[cTail - 1] <= nScenes) cTail--;
      if (c->nColors - cTail < 2) cTail = c->nColors;
    }
    const int nTail = c->nColors - cTail;
    if (nTail > 0) {
      c->tailTable.ensure((size_t)nTail * nScenes + 1);
      CK(cudaMemsetAsync(c->tailTable.p, 0xff, ((size_t)nTail * nScenes + 1) * sizeof(int), c->stream));
      int gT = c->colorStart[cTail];
      LAUNCH(c, k_tail_table, nblk(ng - gT), BLK, gT, ng, cTail, nScenes, c->sgPhase.p, S.sgScene, c->tailTable.p, c->tailTable.p + (size_t)nTail * nScenes);
      int dup = readInt(c, c->tailTable.p + (size_t)nTail * nScenes);
      if (dup != -1) cTail = c->nColors;  // (the flag word starts as -1 like the table)
    }
    const int nTailUse = c->nColors - cTail;
    auto phaseLaunches = [&](int mode, int last) {
      for (int k = 0; k < cTail; k++) {
        int g0 = c->colorStart[k], g1 = c->colorStart[k + 1];
        if (giants[k] > 0) { giantLaunch(mode, g0, giants[k], last); g0 += giants[k]; }
        if (mode == 0) {
          if (hubs) LAUNCH(c, (k_pgs_color<0, true>), nblk(g1 - g0, 128), 128, g0, g1, S, c->dv.p, PP, 0, c->iterState.p);
          else LAUNCH(c, (k_pgs_color<0, false>), nblk(g1 - g0, 128), 128, g0, g1, S, c->dv.p, PP, 0, c->iterState.p);
        } else {
          if (hubs) LAUNCH(c, (k_pgs_color<1, true>), nblk(g1 - g0, 128), 128, g0, g1, S, c->dv.p, PP, last, c->iterState.p);
          else LAUNCH(c, (k_pgs_color<1, false>), nblk(g1 - g0, 128), 128, g0, g1, S, c->dv.p, PP, last, c->iterState.p);
        }
        int r0 = c->colorRunStart[k], r1 = c->colorRunStart[k + 1];
        if (r1 > r0) LAUNCH(c, k_hub_reduce, nblk((long long)(r1 - r0) * 32, 128), 128, r0, r1, HR, S, c->dv.p, c->iterState.p, mode, mode ? PP.check : 0);
        if (mode == 1 && !sweep) c->solveLaunches++;
      }
      if (nTailUse > 0) {
        if (mode == 0) LAUNCH(c, k_pgs_tail<0>, nblk(nScenes, 128), 128, nTailUse, c->tailTable.p, S, c->dv.p, PP, 0, c->iterState.p, 0);
        else LAUNCH(c, k_pgs_tail<1>, nblk(nScenes, 128), 128, nTailUse, c->tailTable.p, S, c->dv.p, PP, last, c->iterState.p, 1);
        if (mode == 1 && !sweep) c->solveLaunches++;
      } else if (mode == 1) {
        LAUNCH(c, k_iter_end, nblk(nScenes), BLK, S, PP, c->iterState.p);
      }
    };
    phaseLaunches(0, 0);
    for (int it = 0; it < iterations; it++) phaseLaunches(1, it == iterations - 1);
  }
  CK(cudaEventRecord(c->ev[sweep ? 9 : 11], c->stream));
  if (!sweep) CK(cudaMemsetAsync(c->bp.nActive.p, 0, (nExt + 1) * sizeof(int), c->stream));
  LAUNCH(c, k_post_solve, nblk(nSolve), BLK, nSolve, S, c->cur.bpc.p, c->cur.lam.p, c->cur.state.p, sweep ? 0 : 1,
         sweep ? (int*)nullptr : c->bp.nActive.p, c->icon.lam.p, c->icon.state.p);
  if (c->recordOrders) {  // tests: keep the Gauss-Seidel sequence for replay on the CPU oracle
    std::vector<int>& dst = sweep ? c->orderSweep : post ? c->orderPost : c->orderFull;
    dst.resize(nSolve);
    if (nSolve) CK(cudaMemcpyAsync(dst.data(), c->scSrc.p, nSolve * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  }
  if (post) {
    CK(cudaStreamSynchronize(c->stream));
  } else if (!sweep) {
    LAUNCH(c, k_row_updates, nblk(ng), BLK, ng, S, nScenes, c->iterState.p);
    unsigned long long st[5];
    readBack(c, st, c->iterState.p, sizeof(st) / sizeof(int));
    c->T.pgs_iterations = (int)st[2];
    c->T.pgs_colors = c->nColors;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev[10], c->ev[11]));
    c->T.pgs_kernel_time = ms * 1e-3;
    c->rowUpdates += 3.0 * (double)st[4];
    c->solveSeconds += c->T.pgs_kernel_time;
  } else {
    CK(cudaStreamSynchronize(c->stream));
  }
}

// ------------------------------------------------------------------------------------------------
// collections: membership CSR + mass properties of the changed ones
// ------------------------------------------------------------------------------------------------
static void rebuildMembers(am3d_ctx* c) {
  int nb = c->NB, nc = c->NS - c->NB;
  c->memKey.ensure(nb + 1); c->memKeySorted.ensure(nb + 1); c->memVal.ensure(nb + 1); c->members.ensure(nb + 1);
  c->collCount.ensure(nc + 2); c->collStart.ensure(nc + 2);
  CK(cudaMemsetAsync(c->collCount.p, 0, (nc + 2) * sizeof(int), c->stream));
  LAUNCH(c, k_member_keys, nblk(nb), BLK, nb, nc, c->parent.p, c->memKey.p, c->memVal.p, c->collCount.p);
  sortPairs(c, c->memKey.p, c->memKeySorted.p, c->memVal.p, c->members.p, nb, 0, bitsFor((unsigned long long)nc + 1));
  c->nMergedLeaves = scanTotal(c, c->collCount, c->collStart, nc);
}
static int countAlive(am3d_ctx* c) {
  int nc = c->NS - c->NB;
  c->tmpI0.ensure(nc + 2); c->tmpI1.ensure(nc + 2);
  CK(cudaMemcpyAsync(c->tmpI0.p, c->collAlive.p, nc * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
  return scanTotal(c, c->tmpI0, c->tmpI1, nc);
}
static void recomputeChanged(am3d_ctx* c) {
  int nc = c->NS - c->NB;
  rebuildMembers(c);
  c->tmpI0.ensure(nc + 2); c->tmpI1.ensure(nc + 2); c->changedList.ensure(nc + 2);
  LAUNCH(c, k_changed_flags, nblk(nc), BLK, nc, c->collAlive.p, c->collMode.p, c->tmpI0.p);
  int nch = scanTotal(c, c->tmpI0, c->tmpI1, nc);
  if (nch > 0) {
    LAUNCH(c, k_changed_list, nblk(nc), BLK, nc, c->tmpI0.p, c->tmpI1.p, c->changedList.p);
    LAUNCH(c, k_coll_recompute, nch, CR_THREADS, nc, c->NB, c->changedList.p, c->collMode.p, c->collStart.p, c->collCount.p,
           c->members.p, c->collFlagAcc.p, c->x.p, c->R.p, c->v.p, c->w.p, c->mass.p, c->minv.p, c->mA.p, c->mA0.p, c->jinv.p,
           c->jinv0.p, c->flags.p, c->bbB.p, c->bbCount.p, c->B2CR.p, c->B2Ct.p);
  }
  CK(cudaMemsetAsync(c->collMode.p, 0, nc * sizeof(int), c->stream));
  c->nCollections = countAlive(c);
}
static int freeSlotList(am3d_ctx* c) {
  int nc = c->NS - c->NB;
  c->tmpI2.ensure(nc + 2); c->tmpI3.ensure(nc + 2); c->freeList.ensure(nc + 2);
  LAUNCH(c, k_free_slots, nblk(nc), BLK, nc, c->collAlive.p, c->tmpI2.p);
  int nfree = scanTotal(c, c->tmpI2, c->tmpI3, nc);
  LAUNCH(c, k_free_list, nblk(nc), BLK, nc, c->tmpI2.p, c->tmpI3.p, c->freeList.p);
  return nfree;
}
static void logEvents(am3d_ctx* c, int kind, const unsigned long long* dkeys, int n) {
  if (n <= 0 || !c->recordEvents) return;
  std::vector<unsigned long long> k(n);
  CK(cudaMemcpyAsync(k.data(), dkeys, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; i++) {
    c->events.push_back(c->totalSteps); c->events.push_back(kind);
    c->events.push_back((int)(k[i] >> 24)); c->events.push_back((int)(k[i] & 0xffffff));
  }
}

// Merging.merge (Merging.java:73-163)
static void mergeStep(am3d_ctx* c) {
  const am3d_params& P = c->P;
  if (!P.enable_merging) return;
  int nbp = c->bp.n;
  if (nbp == 0) return;
  int ns = c->NS, nb = c->NB, nc = ns - nb;
  c->uf.ensure(ns + 1); c->mflag.ensure(nbp + 2);
  LAUNCH(c, k_uf_init, nblk(ns), BLK, ns, c->uf.p);
  CK(cudaMemsetAsync(c->counters.p + 4, 0, sizeof(int), c->stream));
  MergeParamsD MP{P.merge_pinned, P.merge_stable_contact, P.merge_let_it_breathe, P.step_accum_merging, P.threshold_merge, P.threshold_breath};
  LAUNCH(c, k_merge_flag, nblk(nbp, 128), 128, nbp, c->bp.alive.p, c->bp.b1.p, c->bp.b2.p, c->bp.start.p, c->bp.count.p, c->cur.lam.p,
         c->cur.viol.p, c->cur.prevViol.p, c->flags.p, c->parent.p, c->bp.metricHist.p, c->bp.stateHist.p, c->bp.nMetric.p,
         c->bp.nState.p, MP, c->mflag.p, c->uf.p, c->counters.p + 4);
  int nFlagged = readInt(c, c->counters.p + 4);
  if (nFlagged == 0) return;
  CK(cudaEventRecord(c->ev[22], c->stream));  // everything after the conditions = building collections (Merging.java mergingBuildTime)
  c->mergeBuildTimed = true;
  c->mergingEvent = true;
  LAUNCH(c, k_uf_flatten, nblk(ns), BLK, ns, c->uf.p);
  c->compEnt.ensure(ns + 2); c->compBest.ensure(ns + 2); c->needNew.ensure(ns + 2); c->newScan.ensure(ns + 2); c->target.ensure(ns + 2);
  CK(cudaMemsetAsync(c->compEnt.p, 0, ns * sizeof(int), c->stream));
  CK(cudaMemsetAsync(c->compBest.p, 0, ns * sizeof(unsigned long long), c->stream));
  LAUNCH(c, k_merge_census, nblk(ns), BLK, ns, nb, c->collAlive.p, c->parent.p, c->uf.p, c->collCount.p, c->stamp.p, c->compEnt.p, c->compBest.p);
  {  // replay the reference's visiting sequence per component to find the surviving collection (k_mseq_run)
    c->msPar.ensure(ns + 2); c->msSize.ensure(ns + 2); c->msIdent.ensure(ns + 2); c->msSurvivor.ensure(ns + 2);
    c->msKey.ensure(nFlagged + 2); c->msKeySorted.ensure(nFlagged + 2); c->msVal.ensure(nFlagged + 2); c->msValSorted.ensure(nFlagged + 2);
    c->msHead.ensure(nFlagged + 2); c->msScan.ensure(nFlagged + 2); c->msSeg.ensure(nFlagged + 2);
    c->tmpI0.ensure(nbp + 2);
    LAUNCH(c, k_mseq_init, nblk(ns), BLK, ns, nb, c->collAlive.p, c->collCount.p, c->msPar.p, c->msSize.p, c->msIdent.p, c->msSurvivor.p);
    c->msList.ensure(nFlagged + 2); c->msLKey.ensure(nFlagged + 2);
    scanTotal(c, c->mflag, c->tmpI0, nbp);
    LAUNCH(c, k_mseq_list, nblk(nbp), BLK, nbp, c->mflag.p, c->tmpI0.p, c->bp.key.p, c->msList.p, c->msLKey.p);
    if (c->bpTail) {  // pairs an unmerge handed back this step sit at the end of the table: restore ascending (lo, hi)
      c->msList2.ensure(nFlagged + 2); c->msLKey2.ensure(nFlagged + 2);
      sortPairs(c, c->msLKey.p, c->msLKey2.p, c->msList.p, c->msList2.p, nFlagged, 0, 48);
      std::swap(c->msList, c->msList2);
    }
    LAUNCH(c, k_mseq_keys, nblk(nFlagged), BLK, nFlagged, c->msList.p, c->bp.b1.p, c->parent.p, c->uf.p, c->msKey.p, c->msVal.p);
    sortPairs(c, c->msKey.p, c->msKeySorted.p, c->msVal.p, c->msValSorted.p, nFlagged, 0, bitsFor((unsigned long long)ns));
    LAUNCH(c, k_seg_heads, nblk(nFlagged), BLK, nFlagged, c->msKeySorted.p, c->msHead.p);
    int nseg = scanTotal(c, c->msHead, c->msScan, nFlagged);
    LAUNCH(c, k_seg_fill, nblk(nFlagged), BLK, nFlagged, c->msHead.p, c->msScan.p, c->msSeg.p);
    LAUNCH(c, k_mseq_run, nblk(nseg, 64), 64, nseg, c->msSeg.p, c->msKeySorted.p, c->msValSorted.p, c->msList.p, c->bp.b1.p, c->bp.b2.p, c->parent.p,
           c->mergeExactMax, c->msPar.p, c->msSize.p, c->msIdent.p, c->msSurvivor.p);
  }
  LAUNCH(c, k_merge_neednew, nblk(ns), BLK, ns, c->uf.p, c->compEnt.p, c->compBest.p, c->msSurvivor.p, c->needNew.p);
  int nNew = scanTotal(c, c->needNew, c->newScan, ns);
  int nfree = freeSlotList(c);
  if (nNew > nfree) throw AmError(AM3D_ECAPACITY, "out of collection slots");
  LAUNCH(c, k_merge_target, nblk(ns), BLK, ns, nb, c->uf.p, c->compEnt.p, c->compBest.p, c->needNew.p, c->newScan.p, c->freeList.p,
         c->stamp.p, c->collAlive.p, c->msSurvivor.p, c->target.p);
  LAUNCH(c, k_merge_target2, nblk(nc), BLK, nc, nb, c->collAlive.p, c->uf.p, c->collCount.p, c->stamp.p, c->compBest.p, c->target.p);
  CK(cudaMemsetAsync(c->collFlagAcc.p, 0, nc * sizeof(int), c->stream));
  LAUNCH(c, k_merge_newcolls, nblk(ns), BLK, ns, nb, c->uf.p, c->compEnt.p, c->target.p, c->needNew.p, c->newScan.p, c->msSurvivor.p, nFlagged,
         c->collAlive.p, c->flags.p, c->stamp.p, c->nextStamp, c->collMode.p, c->metricCount.p);
  LAUNCH(c, k_merge_apply, nblk(ns), BLK, ns, nb, c->collAlive.p, c->parent.p, c->uf.p, c->compEnt.p, c->target.p, c->needNew.p,
         c->newScan.p, c->flags.p, c->stamp.p, c->nextStamp, c->collMode.p, c->collFlagAcc.p, c->metricCount.p);
  c->nextStamp += (long long)nFlagged + nNew;
  recomputeChanged(c);
  // every live external pair whose two bodies now share a collection becomes internal
  c->tmpI0.ensure(nbp + 2); c->tmpI1.ensure(nbp + 2); c->tmpI2.ensure(nbp + 2); c->tmpI3.ensure(nbp + 2);
  LAUNCH(c, k_int_flag, nblk(nbp), BLK, nbp, c->bp.alive.p, c->bp.b1.p, c->bp.b2.p, c->parent.p, c->bp.nActive.p, c->tmpI0.p, c->tmpI2.p);
  int nIntB = scanTotal(c, c->tmpI0, c->tmpI1, nbp);
  int nIntC = scanTotal(c, c->tmpI2, c->tmpI3, nbp);
  if (nIntB > 0) {
    c->ibp.ensureKeep(c->ibp.n + nIntB + 1, c->ibp.n, c->stream);
    c->ibpCut.ensure(c->ibp.n + nIntB + 1, true, c->stream);
    c->icon.ensureKeep(c->icon.n + nIntC + 1, c->icon.n, c->stream);
    LAUNCH(c, k_int_copy, nblk(nbp, 128), 128, nbp, c->tmpI0.p, c->tmpI1.p, c->tmpI3.p, c->bp.key.p, c->bp.b1.p, c->bp.b2.p, c->bp.start.p,
           c->bp.count.p, contactPtrs(c->cur), c->ibp.n, c->icon.n, c->ibp.key.p, c->ibp.b1.p, c->ibp.b2.p, c->ibp.start.p, c->ibp.count.p,
           c->ibp.alive.p, c->ibp.nMetric.p, c->ibpCut.p, contactPtrs(c->icon));
    logEvents(c, 0, c->ibp.key.p + c->ibp.n, nIntB);
    c->ibp.n += nIntB;
    c->icon.n += nIntC;
  }
}

// Merging.unmerge (Merging.java:215-273); returns true if anything was split
static bool unmergeStep(am3d_ctx* c, double dt) {
  const am3d_params& P = c->P;
  if (!P.enable_unmerging) return false;
  if (!P.unmerge_relative_motion && !P.unmerge_normal && !P.unmerge_friction) return false;
  int nib = c->ibp.n;
  if (nib == 0 || c->nCollections == 0) return false;
  int ns = c->NS, nb = c->NB, nc = ns - nb;
  c->ibpCut.ensure(nib + 1, true, c->stream);
  c->collCuts.ensure(nc + 2); c->collNComp.ensure(nc + 2); c->collKeeps.ensure(nc + 2);
  CK(cudaMemsetAsync(c->collCuts.p, 0, nc * sizeof(int), c->stream));
  CK(cudaMemsetAsync(c->counters.p + 5, 0, sizeof(int), c->stream));
  LAUNCH(c, k_unm_flag, nblk(nib, 128), 128, nib, c->ibp.alive.p, c->ibp.b1.p, c->ibp.b2.p, c->ibp.start.p, c->ibp.count.p, c->icon.state.p,
         c->parent.p, c->flags.p, c->x.p, c->R.p, c->v.p, c->w.p, c->bbB.p, c->bbCount.p, nb, P.threshold_unmerge, P.step_accum_unmerging,
         P.unmerge_normal, P.unmerge_friction, P.metric_position_level ? dt : 0.0, c->ibp.nMetric.p, c->ibpCut.p, c->collCuts.p, c->counters.p + 5);
  int nCuts = readInt(c, c->counters.p + 5);
  if (nCuts == 0) return false;
  CK(cudaEventRecord(c->ev[23], c->stream));  // unmergingBuildTime: splitting the collections
  c->unmergeBuildTimed = true;
  c->uf.ensure(ns + 1);
  LAUNCH(c, k_uf_init, nblk(nb), BLK, nb, c->uf.p);
  LAUNCH(c, k_unm_union, nblk(nib), BLK, nib, c->ibp.alive.p, c->ibpCut.p, c->ibp.b1.p, c->ibp.b2.p, c->parent.p, c->collCuts.p, nb, c->uf.p);
  LAUNCH(c, k_uf_flatten, nblk(nb), BLK, nb, c->uf.p);
  c->compEnt.ensure(ns + 2); c->needNew.ensure(ns + 2); c->newScan.ensure(ns + 2); c->target.ensure(ns + 2);
  c->leavesFlag.ensure(nb + 2); c->leavesScan.ensure(nb + 2); c->freedFlag.ensure(nb + 2);
  CK(cudaMemsetAsync(c->compEnt.p, 0, nb * sizeof(int), c->stream));
  CK(cudaMemsetAsync(c->collNComp.p, 0, nc * sizeof(int), c->stream));
  CK(cudaMemsetAsync(c->collKeeps.p, 0, nc * sizeof(int), c->stream));
  LAUNCH(c, k_unm_census, nblk(nb), BLK, nb, c->parent.p, c->collCuts.p, c->uf.p, c->compEnt.p, c->collNComp.p);
  LAUNCH(c, k_unm_roots, nblk(nb), BLK, nb, c->parent.p, c->collCuts.p, c->uf.p, c->compEnt.p, c->collNComp.p, c->collCount.p,
         c->leavesFlag.p, c->needNew.p, c->freedFlag.p, c->collKeeps.p);
  int nLeaving = scanTotal(c, c->leavesFlag, c->leavesScan, nb);
  bool split = nLeaving > 0;
  if (split) {
    c->mergingEvent = true;
    int nNew = scanTotal(c, c->needNew, c->newScan, nb);
    int nfree = freeSlotList(c);
    if (nNew > nfree) throw AmError(AM3D_ECAPACITY, "out of collection slots");
    CK(cudaMemsetAsync(c->collFlagAcc.p, 0, nc * sizeof(int), c->stream));
    // list position of the leaving pieces (Merging.java:269: bodies.addAll(additionQueue))
    c->grpKey.ensure(nb + 2); c->grpKeySorted.ensure(nb + 2); c->tmpI0.ensure(nb + 2); c->tmpI1.ensure(nb + 2);
    LAUNCH(c, k_unm_rank_keys, nblk(nb), BLK, nb, c->leavesFlag.p, c->parent.p, c->stamp.p, c->grpKey.p, c->tmpI0.p);
    sortPairs(c, c->grpKey.p, c->grpKeySorted.p, c->tmpI0.p, c->tmpI1.p, nb, 0, 64);
    LAUNCH(c, k_unm_rank_scatter, nblk(nLeaving), BLK, nLeaving, c->tmpI1.p, c->leavesScan.p);
    LAUNCH(c, k_unm_apply, nblk(nb), BLK, nb, c->parent.p, c->collCuts.p, c->uf.p, c->leavesFlag.p, c->needNew.p, c->newScan.p, c->freeList.p,
           c->leavesScan.p, c->x.p, c->v.p, c->w.p, c->dv.p, c->stamp.p, c->nextStamp, c->collAlive.p, c->collMode.p, c->flags.p,
           c->metricCount.p);
    c->nextStamp += nLeaving;
    LAUNCH(c, k_unm_retire, nblk(nc), BLK, nc, c->collCuts.p, c->collNComp.p, c->collKeeps.p, c->collAlive.p, c->collMode.p);
    recomputeChanged(c);
  }
  // cut pairs: reconnect or hand back to the external set together with their contacts (Merging.java:343-360)
  c->tmpI0.ensure(nib + 2); c->tmpI1.ensure(nib + 2); c->tmpI2.ensure(nib + 2); c->tmpI3.ensure(nib + 2);
  LAUNCH(c, k_unm_reext_flag, nblk(nib), BLK, nib, c->ibp.alive.p, c->ibpCut.p, c->ibp.b1.p, c->ibp.b2.p, c->ibp.count.p, c->parent.p,
         c->tmpI0.p, c->tmpI2.p);
  int nExtB = scanTotal(c, c->tmpI0, c->tmpI1, nib);
  int nExtC = scanTotal(c, c->tmpI2, c->tmpI3, nib);
  if (nExtB > 0) {
    c->bp.ensureKeep(c->bp.n + nExtB + 1, c->bp.n, c->stream);
    c->cur.ensureKeep(c->cur.n + nExtC + 1, c->cur.n, c->stream);
    LAUNCH(c, k_unm_reext_copy, nblk(nib, 128), 128, nib, c->tmpI0.p, c->tmpI1.p, c->tmpI3.p, c->ibp.key.p, c->ibp.b1.p, c->ibp.b2.p,
           c->ibp.start.p, c->ibp.count.p, c->ibp.alive.p, contactPtrs(c->icon), c->bp.n, c->cur.n, c->bp.key.p, c->bp.b1.p, c->bp.b2.p,
           c->bp.start.p, c->bp.count.p, c->bp.alive.p, c->bp.nActive.p, c->bp.nMetric.p, c->bp.nState.p, contactPtrs(c->cur), c->x.p, c->R.p);
    logEvents(c, 1, c->bp.key.p + c->bp.n, nExtB);
    c->bp.n += nExtB;
    c->cur.n += nExtC;
    c->bpTail = true;
    // compact the internal tables
    LAUNCH(c, k_ibp_compact_flag, nblk(nib), BLK, nib, c->ibp.alive.p, c->ibp.count.p, c->tmpI0.p, c->tmpI2.p);
    int keepB = scanTotal(c, c->tmpI0, c->tmpI1, nib);
    int keepC = scanTotal(c, c->tmpI2, c->tmpI3, nib);
    c->ibp2.ensure(keepB + 1); c->icon2.ensure(keepC + 1); c->ibpCut2.ensure(keepB + 1);
    LAUNCH(c, k_ibp_compact, nblk(nib, 128), 128, nib, c->tmpI0.p, c->tmpI1.p, c->tmpI3.p, c->ibp.key.p, c->ibp.b1.p, c->ibp.b2.p, c->ibp.start.p,
           c->ibp.count.p, c->ibp.nMetric.p, contactPtrs(c->icon), c->ibp2.key.p, c->ibp2.b1.p, c->ibp2.b2.p, c->ibp2.start.p, c->ibp2.count.p,
           c->ibp2.alive.p, c->ibp2.nMetric.p, c->ibpCut2.p, contactPtrs(c->icon2));
    std::swap(c->ibp, c->ibp2);
    std::swap(c->icon, c->icon2);
    std::swap(c->ibpCut, c->ibpCut2);
    c->ibp.n = keepB;
    c->icon.n = keepC;
  }
  return split;
}
