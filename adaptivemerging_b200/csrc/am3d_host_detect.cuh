// Host orchestration of detection, body-pair bookkeeping and warm start.
#pragma once
#include "am3d_host_util.cuh"
#include "am3d_sort.cuh"
#include "am3d_detect.cuh"
#include "am3d_step.cuh"
static void detect(am3d_ctx* c) {
  int nsh = c->NSH;
  std::swap(c->cur, c->prev);  // ContactPool.swapPools (ContactPool.java:60-65): last step's contacts stay readable
  LAUNCH(c, k_shape_update, nblk(nsh), BLK, nsh, c->shType.p, c->shBody.p, c->shRoot.p, c->shRadius.p, c->shSize.p, c->shLR.p, c->shLt.p,
         c->btype.p, c->x.p, c->R.p, c->ndC.p, c->ndR.p, c->shX.p, c->shR.p, c->shBoundC.p, c->shBoundR.p, c->shBoundH.p);
  double inv = 1.0 / c->cellSize;
  if (c->pairKey.cap == 0) {
    size_t cap = (size_t)nsh * 8 + 1024;
    c->pairKey.ensure(cap); c->pairVal.ensure(cap); c->pairKeySorted.ensure(cap); c->pairValSorted.ensure(cap);
  }
  if (c->nSmall > 0) {
    LAUNCH(c, k_cell_keys, nblk(c->nSmall), BLK, c->nSmall, c->smallList.p, c->shBody.p, c->scene.p, c->shBoundC.p, inv,
           c->cellKey.p, c->cellVal.p);
    int endBit = std::min(64, 42 + bitsFor((unsigned long long)c->H.nscenes));
    sortPairs(c, c->cellKey.p, c->cellKeySorted.p, c->cellVal.p, c->cellValSorted.p, c->nSmall, 0, endBit);
  }
  int np = 0;
  for (int attempt = 0; attempt < 3; attempt++) {
    CK(cudaMemsetAsync(c->counters.p, 0, sizeof(int), c->stream));
    PairCtx PC{c->shBody.p, c->bShapeFirst.p, c->parent.p, c->flags.p, c->scene.p, c->stamp.p, c->shBoundC.p, c->shBoundR.p, c->shBoundH.p,
               c->pairKey.p, c->pairVal.p, c->counters.p, (int)std::min<size_t>(c->pairKey.cap, 0x7fffffff),
               bitsFor((unsigned long long)c->NB), c->haveComposites ? 16 : 0};
    if (c->nSmall > 0)
      LAUNCH(c, k_pairs_grid, nblk(c->nSmall), BLK, c->nSmall, c->cellKeySorted.p, c->cellValSorted.p, inv, PC);
    if (c->nLarge > 0 || c->nPlanes > 0)
      LAUNCH(c, k_pairs_special, nblk(nsh), BLK, nsh, c->shType.p, c->shLarge.p, c->largeStart.p, c->largeList.p, c->planeStart.p,
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
    int endBit = 2 * bitsFor((unsigned long long)c->NB) + (c->haveComposites ? 16 : 0);
    sortPairs(c, c->pairKey.p, c->pairKeySorted.p, c->pairVal.p, c->pairValSorted.p, np, 0, endBit);
    c->pairType.ensure(np + 1); c->pairCap.ensure(np + 1); c->pairSlot.ensure(np + 1); c->pairCount.ensure(np + 1); c->pairOut.ensure(np + 1);
    c->treeList.ensure(np + 1);
    CK(cudaMemsetAsync(c->counters.p + 7, 0, sizeof(int), c->stream));
    LAUNCH(c, k_pair_classify, nblk(np), BLK, np, c->pairValSorted.p, c->shType.p, c->pairType.p, c->pairCap.p, c->treeList.p,
           c->counters.p + 7);
    int treeGrid = std::min(nblk(np, WARPS_PER_BLOCK), 148 * 16);
    TreeCtx TC{c->shType.p, c->shRoot.p, c->shSize.p, c->shRadius.p, c->shP.p, c->shX.p, c->shR.p, c->ndC.p, c->ndR.p, c->ndFirst.p, c->ndCount.p};
    HitOut HO{c->hitPos.p, c->hitNrm.p, c->hitViol.p, c->hitMeta.p};
    bool haveTrees = c->NN > 0;
    CK(cudaMemsetAsync(c->counters.p + 1, 0, sizeof(int), c->stream));
    CK(cudaEventRecord(c->ev[16], c->stream));
    // sphere-tree pairs: count pass.  A tree x tree pair whose frontier grows past SPLIT_TARGET node pairs hands them to
    // the task list (one warp per node pair, k_tree_tasks); the pair's contacts are then [its own prefix walk][task 0][task 1]...
    TreeTasks TT{};
    int nTasks = 0;
    if (haveTrees) {
      c->treePairStart.ensure(np + 1); c->treePairN.ensure(np + 1); c->treePairPrefix.ensure(np + 1);
      if (c->taskVal.cap == 0) { c->taskVal.ensure(65536); c->taskPair.ensure(65536); }
      for (int attempt = 0; attempt < 3; attempt++) {
        CK(cudaMemsetAsync(c->counters.p + 10, 0, sizeof(int), c->stream));
        TT = TreeTasks{c->taskVal.p, c->taskPair.p, nullptr, nullptr, c->counters.p + 10, (int)std::min<size_t>(c->taskVal.cap, 0x7fffffff), c->treeSplit ? SPLIT_TARGET : 0x7fffffff,
                       c->treePairStart.p, c->treePairN.p, c->treePairPrefix.p};
        LAUNCH(c, k_narrow_tree<false>, treeGrid, WARPS_PER_BLOCK * 32, c->treeList.p, c->counters.p + 7, c->pairValSorted.p, c->pairType.p,
               c->pairSlot.p, TC, HO, c->pairCap.p, c->counters.p + 1, TT);
        nTasks = readInt(c, c->counters.p + 10);
        if ((size_t)nTasks <= c->taskVal.cap) break;
        size_t cap = (size_t)nTasks + nTasks / 4 + 1024;
        c->taskVal.ensure(cap); c->taskPair.ensure(cap);
      }
      if (nTasks > 0) {
        c->taskCount.ensure(nTasks + 2); c->taskPrefix.ensure(nTasks + 2);
        TT.count = c->taskCount.p;
        int taskGrid = std::min(nblk(nTasks, WARPS_PER_BLOCK), 148 * 32);
        LAUNCH(c, k_tree_tasks<false>, taskGrid, WARPS_PER_BLOCK * 32, nTasks, c->pairValSorted.p, c->pairSlot.p, TC, HO, c->counters.p + 1, TT);
        CK(cudaMemsetAsync(c->taskCount.p + nTasks, 0, sizeof(int), c->stream));
        exclusiveSum(c, c->taskCount.p, c->taskPrefix.p, nTasks + 1);
        TT.prefix = c->taskPrefix.p;
        LAUNCH(c, k_tree_paircap, nblk(np), BLK, c->treeList.p, c->counters.p + 7, TT, c->pairCap.p);
      }
    }
    CK(cudaEventRecord(c->ev[17], c->stream));
    int nslots = scanTotal(c, c->pairCap, c->pairSlot, np);
    c->nSlots = nslots;
    c->hitPos.ensure(3 * (size_t)nslots + 3); c->hitNrm.ensure(3 * (size_t)nslots + 3); c->hitViol.ensure((size_t)nslots + 1);
    c->hitMeta.ensure(4 * (size_t)nslots + 4);
    HO = HitOut{c->hitPos.p, c->hitNrm.p, c->hitViol.p, c->hitMeta.p};
    CK(cudaMemsetAsync(c->pairCount.p, 0, (np + 1) * sizeof(int), c->stream));
    CK(cudaEventRecord(c->ev[18], c->stream));
    LAUNCH(c, k_narrow_box, nblk(np, 128), 128, np, c->pairValSorted.p, c->pairType.p, c->pairSlot.p, c->shSize.p, c->shRadius.p,
           c->shX.p, c->shR.p, HO, c->pairCount.p);
    if (haveTrees) {
      LAUNCH(c, k_narrow_tree<true>, treeGrid, WARPS_PER_BLOCK * 32, c->treeList.p, c->counters.p + 7, c->pairValSorted.p, c->pairType.p,
             c->pairSlot.p, TC, HO, c->pairCount.p, c->counters.p + 1, TT);
      if (nTasks > 0) {
        int taskGrid = std::min(nblk(nTasks, WARPS_PER_BLOCK), 148 * 32);
        LAUNCH(c, k_tree_tasks<true>, taskGrid, WARPS_PER_BLOCK * 32, nTasks, c->pairValSorted.p, c->pairSlot.p, TC, HO, c->counters.p + 1, TT);
      }
    }
    CK(cudaEventRecord(c->ev[19], c->stream));
    c->narrowTimed = true;
    nc = scanTotal(c, c->pairCount, c->pairOut, np);
    if (haveTrees && readInt(c, c->counters.p + 1)) throw AmError(AM3D_ECAPACITY, "sphere-tree traversal stack overflow");
    c->cur.ensure(nc + 1);
    ContactOut CO{c->cur.b1.p, c->cur.b2.p, c->cur.s1.p, c->cur.s2.p, c->cur.bv1.p, c->cur.bv2.p, c->cur.info.p, c->cur.leaf.p,
                  c->cur.state.p, c->cur.isNew.p, c->cur.key0.p, c->cur.key1.p, c->cur.pW.p, c->cur.nW.p, c->cur.t1W.p,
                  c->cur.t2W.p, c->cur.pB1.p, c->cur.nB1.p, c->cur.t1B1.p, c->cur.t2B1.p, c->cur.viol.p, c->cur.prevViol.p,
                  c->cur.lam.p, c->cur.lamWarm.p};
    if (nc > 0) {
      c->tmpI3.ensure(nc + 1);
      LAUNCH(c, k_contact_owner, nblk(np), BLK, np, c->pairCount.p, c->pairOut.p, c->tmpI3.p);
      LAUNCH(c, k_contact_set, nblk(nc, 128), 128, nc, c->tmpI3.p, c->bShapeFirst.p, c->pairValSorted.p, c->pairSlot.p, c->pairOut.p,
             c->shBody.p, c->x.p, c->R.p, c->hitPos.p, c->hitNrm.p, c->hitViol.p, c->hitMeta.p, CO);
    }
  }
  c->cur.n = nc;
  c->cur.nSorted = nc;
  c->bpTail = false;
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
    c->bpSlot ^= 1;  // [8 + slot]: largest contact count of a pair in THIS detection, [8 + (slot ^ 1)]: in the one before
    CK(cudaMemsetAsync(c->counters.p + 8 + c->bpSlot, 0, sizeof(int), c->stream));
    LAUNCH(c, k_bpc_match, nblk(nbp), BLK, nbp, nc, c->bp.key.p, c->bp.start.p, c->bp.count.p, c->bp.b1.p, c->bp.b2.p,
           c->bp.nActive.p, c->bp.metricHist.p, c->bp.stateHist.p, c->bp.nMetric.p, c->bp.nState.p, c->bp.alive.p,
           c->bpPrev.n, c->bpPrev.key.p, c->bpPrev.b1.p, c->bpPrev.b2.p, c->bpPrev.metricHist.p, c->bpPrev.stateHist.p,
           c->bpPrev.nMetric.p, c->bpPrev.nState.p, c->counters.p + 8 + c->bpSlot);
  }
  c->bp.n = nbp;
}

// postStab: warmStart(true) of postStabilization keeps the violation the donor itself inherited (CollisionProcessor.java:569-572)
static void warmStart(am3d_ctx* c, bool postStab = false) {
  int nbp = c->bp.n;
  if (nbp == 0) return;
  int nt = c->prev.n - c->prev.nSorted;
  if (nt > 0) {  // index the out-of-order tail of last step's contact list (appended by an unmerge)
    c->tailKey.ensure(nt + 1); c->tailKeySorted.ensure(nt + 1); c->tailVal.ensure(nt + 1); c->tailIdx.ensure(nt + 1);
    LAUNCH(c, k_tail_keys, nblk(nt), BLK, nt, c->prev.nSorted, c->prev.key0.p, c->tailKey.p, c->tailVal.p);
    sortPairs(c, c->tailKey.p, c->tailKeySorted.p, c->tailVal.p, c->tailIdx.p, nt, 0, 64);
  }
  c->wsPairSlow.ensure(nbp + 1); c->wsMatch.ensure(c->cur.n + 1);
  CK(cudaMemsetAsync(c->wsPairSlow.p, 0, nbp * sizeof(int), c->stream));
  int ns = c->prev.nSorted;
  // a pair of sphere trees can hold 10^4+ contacts: index last step's contacts when it had such pairs
  int useIdx = c->NN > 0 && ns > 0 && readInt(c, c->counters.p + 8 + (c->bpSlot ^ 1)) > 64;
  if (useIdx) {  // stable sort of the canonical entries by key1, then by key0
    c->wsKa.ensure(ns + 1); c->wsKb.ensure(ns + 1); c->wsK0s.ensure(ns + 1); c->wsK1s.ensure(ns + 1);
    c->wsIa.ensure(ns + 1); c->wsIb.ensure(ns + 1); c->wsIdx.ensure(ns + 1);
    LAUNCH(c, k_iota, nblk(ns), BLK, ns, c->wsIa.p);
    sortPairs(c, c->prev.key1.p, c->wsKa.p, c->wsIa.p, c->wsIb.p, ns, 0, 64);
    LAUNCH(c, k_gather_u64, nblk(ns), BLK, ns, c->wsIb.p, c->prev.key0.p, c->wsKb.p);
    sortPairs(c, c->wsKb.p, c->wsK0s.p, c->wsIb.p, c->wsIdx.p, ns, 0, 64);
    LAUNCH(c, k_gather_u64, nblk(ns), BLK, ns, c->wsIdx.p, c->prev.key1.p, c->wsK1s.p);
  }
  WarmCtx W{c->cur.b1.p, c->cur.b2.p, c->cur.s1.p, c->cur.s2.p, c->cur.leaf.p, c->cur.key0.p, c->cur.key1.p, c->cur.pB1.p,
            c->cur.lam.p, c->cur.lamWarm.p, c->cur.prevViol.p, c->cur.isNew.p,
            c->prev.n, c->prev.nSorted, c->cur.n, c->tailKeySorted.p, c->tailIdx.p, c->prev.key0.p, c->prev.key1.p, c->prev.b1.p, c->prev.leaf.p, c->prev.pB1.p, postStab ? c->prev.prevViol.p : c->prev.viol.p, c->prev.lam.p,
            useIdx, c->wsK0s.p, c->wsK1s.p, c->wsIdx.p,
            c->btype.p, c->shType.p, c->x.p, c->R.p, c->ndRank.p, c->cur.bpc.p, c->wsPairSlow.p, c->wsMatch.p};
  LAUNCH(c, k_warm_start_plain, nblk(c->cur.n), BLK, c->cur.n, W);
  LAUNCH(c, k_warm_apply, nblk(c->cur.n), BLK, c->cur.n, W);
  LAUNCH(c, k_warm_start, nblk(nbp, 128), 128, nbp, c->bp.start.p, c->bp.count.p, c->bp.b1.p, c->bp.b2.p, W);
}

// ------------------------------------------------------------------------------------------------
// solve
// ------------------------------------------------------------------------------------------------
