// Device-resident state of one am3d context (one GPU, one stream).
//
// Layout in HBM (DESIGN.md "Data layout"): every per-body / per-shape / per-contact quantity is a
// separate flat array (structure of arrays), doubles for all physics, int32 for ids and flags.  Bodies
// live in a "solver body" index space [0,NB) = leaf bodies in XML parse order, [NB, NB+NCcap) =
// RigidCollection slots, so that the PGS kernels address a merged collection exactly like a free body.
#pragma once
#include <chrono>
#include <future>
#include <cstdio>
#include <cstdlib>
#define DVS 8  // deltaV stride in doubles: 6 used, padded so that a body is two aligned 32-byte accesses
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/am3d.h"

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) throw AmError(AM3D_ECUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct AmError {
  int code;
  std::string msg;
  AmError(int c, const std::string& m) : code(c), msg(m) {}
};

// stream of the context whose API call is running on this thread (set by API_BEGIN)
inline cudaStream_t& amCurrentStream() { static thread_local cudaStream_t s = nullptr; return s; }
inline cudaMemPool_t& amCurrentPool() { static thread_local cudaMemPool_t p = nullptr; return p; }  // the context's private pool
inline size_t& amAllocatedBytes() { static size_t b = 0; return b; }  // device bytes handed out by DevBuf (all contexts)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { T* tp = p; size_t tc = cap; p = o.p; cap = o.cap; o.p = tp; o.cap = tc; }
    return *this;
  }
  ~DevBuf() { if (p) cudaFree(p); }
  // Growth goes through the stream-ordered allocator on the calling context's stream: cudaFree / cudaMalloc in the
  // middle of a run cost 10-200 ms each on B200 (measured), cudaMallocAsync / cudaFreeAsync from the retained pool
  // (a private pool per context with release threshold = max, created in am3d_create) cost microseconds and need no device-wide synchronisation.
  void ensure(size_t n, bool keep = false, cudaStream_t st = 0) {
    if (n <= cap) return;
    size_t ncap = n + n / 2 + 256;  // grow geometrically
    cudaStream_t s = st ? st : amCurrentStream();
    T* q = nullptr;
    static const bool traceAlloc = getenv("AM3D_TRACE_ALLOC") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    if (amCurrentPool()) CK(cudaMallocFromPoolAsync(&q, ncap * sizeof(T), amCurrentPool(), s));
    else CK(cudaMallocAsync(&q, ncap * sizeof(T), s));
    if (keep && p && cap) CK(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (p) CK(cudaFreeAsync(p, s));  // ordered after everything already queued on the stream that uses p
    if (traceAlloc)
      fprintf(stderr, "[am3d alloc] %.1f MB (was %.1f MB): %.3f ms\n", ncap * sizeof(T) / 1e6, cap * sizeof(T) / 1e6,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    amAllocatedBytes() += (ncap - cap) * sizeof(T);
    p = q;
    cap = ncap;
  }
  void zero(size_t n, cudaStream_t st) { if (n) CK(cudaMemsetAsync(p, 0, n * sizeof(T), st)); }
  void free_() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  operator T*() const { return p; }
};

// per-contact arrays (one set for the current step, one for the previous step: warm-start source)
struct ContactSet {
  int n = 0;
  int nSorted = 0;  // entries [0,nSorted) are in canonical key order; the rest was appended by an unmerge
  DevBuf<int> b1, b2;        // Contact.body1/body2 (leaf or composite-parent ids)
  DevBuf<int> s1, s2;        // shapes (part = shape - body_shape_first)
  DevBuf<int> bv1, bv2, info, leaf;
  DevBuf<int> bpc;           // index of the body pair this contact belongs to (-1: pinned-pinned)
  DevBuf<int> state, isNew;
  DevBuf<unsigned long long> key0;  // (bodyLo<<40 | bodyHi<<16 | partLo<<8 | partHi)
  DevBuf<unsigned long long> key1;  // (bvLo+2)<<36 | (bvHi+2)<<8 | info      (normalised to lo/hi side)
  DevBuf<double> pW, nW, t1W, t2W;  // world frame at Contact.set time [3]
  DevBuf<double> pB1, nB1, t1B1, t2B1;  // same frame in body1 coordinates [3]
  DevBuf<double> viol, prevViol;
  DevBuf<double> lam, lamWarm;      // [3]
  void ensure(size_t c) {
    b1.ensure(c); b2.ensure(c); s1.ensure(c); s2.ensure(c); bv1.ensure(c); bv2.ensure(c); info.ensure(c); leaf.ensure(c);
    bpc.ensure(c); state.ensure(c); isNew.ensure(c); key0.ensure(c); key1.ensure(c);
    pW.ensure(3 * c); nW.ensure(3 * c); t1W.ensure(3 * c); t2W.ensure(3 * c);
    pB1.ensure(3 * c); nB1.ensure(3 * c); t1B1.ensure(3 * c); t2B1.ensure(3 * c);
    viol.ensure(c); prevViol.ensure(c); lam.ensure(3 * c); lamWarm.ensure(3 * c);
  }
  // grow while keeping the existing entries
  void ensureKeep(size_t c, size_t, cudaStream_t st) {
    if (c > b1.cap) c = c + c / 2;  // amortise; every field is checked on its own (capacities are not proportional)
    b1.ensure(c, true, st); b2.ensure(c, true, st); s1.ensure(c, true, st); s2.ensure(c, true, st); bv1.ensure(c, true, st);
    bv2.ensure(c, true, st); info.ensure(c, true, st); leaf.ensure(c, true, st); bpc.ensure(c, true, st); state.ensure(c, true, st);
    isNew.ensure(c, true, st); key0.ensure(c, true, st); key1.ensure(c, true, st);
    pW.ensure(3 * c, true, st); nW.ensure(3 * c, true, st); t1W.ensure(3 * c, true, st); t2W.ensure(3 * c, true, st);
    pB1.ensure(3 * c, true, st); nB1.ensure(3 * c, true, st); t1B1.ensure(3 * c, true, st); t2B1.ensure(3 * c, true, st);
    viol.ensure(c, true, st); prevViol.ensure(c, true, st); lam.ensure(3 * c, true, st); lamWarm.ensure(3 * c, true, st);
  }
};

// per body-pair arrays (BodyPairContact.java)
struct BpcSet {
  int n = 0;
  DevBuf<unsigned long long> key;  // bodyLo<<32 | bodyHi, ascending
  DevBuf<int> b1, b2;              // orientation at creation
  DevBuf<int> start, count;        // contact range in the ContactSet
  DevBuf<int> nActive;             // contacts with |lambda0| > 1e-14 after the solve
  DevBuf<double> metricHist;       // [4]
  DevBuf<int> stateHist;           // [4]
  DevBuf<int> nMetric, nState;
  DevBuf<int> alive;
  void ensure(size_t c) {
    key.ensure(c); b1.ensure(c); b2.ensure(c); start.ensure(c); count.ensure(c); nActive.ensure(c);
    metricHist.ensure(4 * c); stateHist.ensure(4 * c); nMetric.ensure(c); nState.ensure(c); alive.ensure(c);
  }
  void ensureKeep(size_t c, size_t, cudaStream_t st) {
    if (c > key.cap) c = c + c / 2;
    key.ensure(c, true, st); b1.ensure(c, true, st); b2.ensure(c, true, st); start.ensure(c, true, st); count.ensure(c, true, st);
    nActive.ensure(c, true, st); metricHist.ensure(4 * c, true, st); stateHist.ensure(4 * c, true, st); nMetric.ensure(c, true, st);
    nState.ensure(c, true, st); alive.ensure(c, true, st);
  }
};

struct am3d_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaMemPool_t pool = nullptr;  // private stream-ordered allocation pool (nothing of the process's default pool is reconfigured)
  std::string lastError;
  std::future<int> pending;  // steps handed to the worker thread by am3d_step_async
  int asyncStatus = 0;
  am3d_params P;
  bool haveScene = false;
  int totalSteps = 0;
  bool mergingEvent = false;

  // ---- host copy of the scene (for reset) -------------------------------------------------
  struct HostScene {
    int nb = 0, nsh = 0, nn = 0, nsp = 0, nscenes = 1;
    std::vector<int> body_type, body_flags, body_scene, body_shape_first, body_shape_count, body_bb_count;
    std::vector<double> body_x, body_R, body_v, body_omega, body_mass, body_minv, body_mA0, body_jinv0, body_fric,
        body_rest, body_bbB;
    std::vector<int> shape_type, shape_body, shape_root;
    std::vector<double> shape_size, shape_radius, shape_p, shape_lR, shape_lt;
    std::vector<double> node_c, node_r;
    std::vector<int> node_first, node_count, node_rank;
    std::vector<int> sp_type, sp_b1, sp_b2;
    std::vector<double> sp_pb1, sp_pb2, sp_pw, sp_k, sp_d, sp_l0, sp_ls;
  } H;

  int NB = 0, NS = 0, NSH = 0, NN = 0, NSP = 0;

  // ---- solver bodies [NS] ---------------------------------------------------------------------
  DevBuf<double> x, R, v, w, force, torque, dv, minv, mass, jinv, jinv0, mA, mA0, fric, rest, bbB;
  DevBuf<int> bbCount, flags, scene, parent, btype, collAlive;
  int nCollections = 0;
  int nMergedLeaves = 0;  // leaves that currently belong to a collection
  long long nextStamp = 0;
  DevBuf<long long> stamp;
  DevBuf<double> metricHist;  // [10] ring (ordered oldest..newest)
  DevBuf<int> metricCount, hasExt;
  DevBuf<int> bShapeFirst, bShapeCount;

  // ---- shapes [NSH] ---------------------------------------------------------------------------
  DevBuf<int> shType, shBody, shRoot, shLarge;
  DevBuf<double> shSize, shRadius, shP, shLR, shLt, shX, shR, shBoundC, shBoundR, shBoundH;  // shX/shR: world transform this step
  // ---- sphere-tree nodes [NN] -------------------------------------------------------------------
  DevBuf<double> ndC, ndR;
  DevBuf<int> ndFirst, ndCount, ndRank;
  // ---- springs ------------------------------------------------------------------------------------
  DevBuf<int> spType, spB1, spB2;
  DevBuf<double> spPb1, spPb2, spPw, spK, spD, spL0, spLs;
  DevBuf<int> spBodyStart, spBodyList;  // CSR body -> (spring<<1 | side), in spring order
  int nSpringBodies = 0, nBodyBodySprings = 0;
  DevBuf<int> spBodies;

  // ---- broadphase ---------------------------------------------------------------------------------
  std::vector<int> hSmall, hLarge, hPlanes;
  DevBuf<int> smallList, largeList, planeList, largeStart, planeStart;  // special lists are grouped by scene
  int nSmall = 0, nLarge = 0, nPlanes = 0;
  bool haveComposites = false;  // some body owns more than one shape: pair keys carry the part indices
  double cellSize = 1.0;
  DevBuf<unsigned long long> cellKey, cellKeySorted;
  DevBuf<int> cellVal, cellValSorted;
  DevBuf<unsigned long long> pairKey, pairKeySorted, pairVal, pairValSorted;
  DevBuf<int> counters;  // small device scalars
  DevBuf<unsigned char> cubTemp;
  int nPairs = 0;
  DevBuf<int> pairType, pairCap, pairSlot, pairCount, pairOut;
  // raw hits (slot layout)
  DevBuf<double> hitPos, hitNrm, hitViol;
  DevBuf<int> treeList;  // indices of the candidate pairs that involve a sphere tree
  DevBuf<unsigned long long> taskVal;  // tree x tree pairs split into one task per waiting node pair (k_tree_tasks)
  DevBuf<int> taskPair, taskCount, taskPrefix, treePairStart, treePairN, treePairPrefix;
  DevBuf<int> hitMeta;  // [4]: info, bv1, bv2, leaf
  long long nSlots = 0;

  ContactSet cur, prev;
  BpcSet bp, bpPrev, bpTmp;
  int bpSlot = 0;       // which of counters[8..9] holds this detection's largest pair
  bool bpTail = false;  // bp holds pairs appended by an unmerge (not in key order)

  // ---- merging ------------------------------------------------------------------------------------
  ContactSet icon, icon2;   // internal contacts of collections (RigidCollection.internalContacts), grouped by pair
  BpcSet ibp, ibp2;         // internal body pairs (BodyPairContact.inCollection == true)
  DevBuf<int> ibpCut, ibpCut2;
  DevBuf<double> pokeV, pokeW;
  DevBuf<unsigned char> mouse;  // one MouseState (am3d_step.cuh): mouse spring + impulse of the UI
  DevBuf<int> picked;           // RigidBody.picked per leaf
  bool mouseUsed = false;
  int nDormant = 0;             // bodies of the blob that are not in RigidBodySystem.bodies (AM3D_F_DORMANT)
  DevBuf<unsigned long long> tailKey, tailKeySorted;
  DevBuf<int> tailVal, tailIdx;
  DevBuf<unsigned long long> wsKa, wsKb, wsK0s, wsK1s;  // (key0,key1)-sorted index of last step's contacts
  DevBuf<int> wsIa, wsIb, wsIdx, wsPairSlow, wsMatch;
  DevBuf<double> B2CR, B2Ct;  // RigidBody.transformB2C of the leaves
  DevBuf<int> collCount, collStart, members, memVal, changedList, collMode, collFlagAcc, freeList;
  DevBuf<unsigned int> memKey, memKeySorted;
  DevBuf<int> uf, mflag, compEnt, needNew, newScan, target, collCuts, collNComp, collKeeps, leavesFlag, leavesScan, freedFlag;
  DevBuf<unsigned long long> compBest;
  DevBuf<int> msPar, msSize, msIdent, msSurvivor, msVal, msValSorted, msHead, msScan, msSeg;  // sequential replay of Merging.merge
  DevBuf<unsigned int> msKey, msKeySorted;
  DevBuf<int> msList, msList2;
  DevBuf<unsigned long long> msLKey, msLKey2;
  int mergeExactMax = 16384;  // mergeable pairs per component up to which the reference's visiting sequence is replayed
  DevBuf<int> swB1, swB2, swCount, swStart, swAsleep, tmpI2, tmpI3;
  DevBuf<int> grpLayer, bodyLevel, bfsRound;  // breadth-first layers of the single sweep (getOrganizedContacts)
  DevBuf<int> bodyLocal;   // rank of a leaf body among the bodies of its scene (colour priorities hash scene-local ids)
  DevBuf<int> sgScene;     // per solve group
  DevBuf<int> sceneState;  // per scene: done | moving | iterations (the tolerance exit is taken per scene)
  DevBuf<int> phaseHead, phaseScan, sgPhase, phaseGiants;
  DevBuf<int> chN, chFirst, cgB1, cgB2, cgCount, cgStart, cgLayer, cgLead;  // body pairs cut into chunks of giantChunk contacts
  int giantChunk = 512;   // am3d_set_option("giant_chunk", n): 0 = never split a pair
  int nPairsSolve = 0;    // body pairs of the last solve (nGroups counts chunks)
  // partitions of batched scenes, one thread-block cluster each (k_pgs_cluster)
  int useClusters = 1, maxClusters = 0, nPart = 0, hPartScenes = 0;
  DevBuf<int> partRange, partSceneStart, partRemaining;
  std::vector<int> hPartSceneStart, hPartCount;
  int ownPrimitives = 1;  // am3d_set_option("own_primitives", 0/1): hand-written radix sort / prefix sum (am3d_sort.cuh) or the CUB ones
  DevBuf<unsigned long long> rsKeyTmp, rsValTmp;
  DevBuf<int> rsHist, rsOff, scanSums;
  int useSceneBfs = 1;    // am3d_set_option("scene_bfs", 0/1): breadth-first layers of the single sweep per scene (k_bfs_scenes) for batched contexts
  DevBuf<int> sceneCnt, sceneStart, sceneCursor, sceneList;
  int useTailFusion = 1;  // am3d_set_option("pgs_tail_fusion", 0/1): trailing phases with one group per scene in one launch (k_pgs_tail)
  DevBuf<int> tailTable;
  int treeSplit = 1;      // am3d_set_option("tree_split", 0/1): tree x tree pairs with a large frontier are split into one task per node pair
  int fastRows = 1;       // am3d_set_option("pgs_fast_rows", 0/1): branch-free PGS row update (bit-identical results; 0 = the plain form everywhere)
  int useGiantWarps = 1;  // am3d_set_option("giant_warps", 0/1): groups of >= 65 contacts are solved by a warp (k_pgs_giant)  // (layer, colour) phases of the sorted group list
  int bfsBlocks = 0, colorBlocks = 0;
  DevBuf<int> colorCtl;
  bool orderingTimed = false, mergeBuildTimed = false, unmergeBuildTimed = false;
  std::vector<int> events;  // (step, kind, bodyLo, bodyHi) quadruples
  bool recordEvents = true; // am3d_set_option("record_events", 0): long batched runs (one device->host copy per merge step saved)
  bool recordOrders = false;
  std::vector<int> orderFull, orderSweep, orderPost;
  std::vector<am3d_contact> orderFullKeys, orderSweepKeys, orderPostKeys;
  bool lastSolveSweep = false;
  int lastSolveN = 0;

  // ---- solver (colour order) ----------------------------------------------------------------------
  DevBuf<int> grpColor, grpOrder, grpSb1, grpSb2, grpPos, grpVal, colorHist, tmpI0, tmpI1;
  DevBuf<unsigned long long> grpKey, grpKeySorted;
  DevBuf<unsigned long long> grpPrio, bodyBest, bodyMask;
  DevBuf<int> grpList, grpList2;
  DevBuf<int> sgB1, sgB2, sgStart, sgCount, sgFlags, sgBpc;  // per solve group (colour-major)
  DevBuf<double> sgMass;                            // [20] minv1,jinv1,minv2,jinv2
  DevBuf<double> sgMu;
  DevBuf<double> scP;                               // per contact in solve order: [24] n t1 t2 | r1 r2 | b | D | lambda
  DevBuf<double> hubDelta;                          // [12] per group
  DevBuf<int> grpDegree, grpHubMask, hubN, hubScan, hubSlot, hubSlotSorted, hubHead, hubRunStart, hubRunBody, hubRunColor, dColorRunStart;
  DevBuf<unsigned long long> hubKey, hubKeySorted;
  std::vector<int> colorRunStart;                   // host: first hub run of each phase
  int nHubRuns = 0, nHubEntries = 0;
  int hubMin = 64;                                  // degree (in body pairs) from which a body is treated as a hub; 0 = never
  DevBuf<int> scSrc;                                // solve-order -> canonical contact index
  DevBuf<int> scState;
  std::vector<int> colorStart;                      // host copy, ncolors+1
  DevBuf<int> dColorStart;
  int nColors = 0, nGroups = 0;
  DevBuf<unsigned long long> iterState;             // [0]=max bits, [1]=done, [2]=iters executed

  // ---- timing --------------------------------------------------------------------------------------
  am3d_timings T;
  cudaEvent_t ev[24];
  cudaEvent_t evPoke = nullptr, evMain = nullptr;  // velocity pokes uploaded on the copy stream (am3d_add_velocities)
  bool pokesPending = false;
  int* mappedHost = nullptr;   // mapped pinned memory for the small read-backs of a step (readBack in am3d_host_util.cuh)
  int* mappedDev = nullptr;
  cudaStream_t upStream = nullptr;
  cudaStream_t copyStream = nullptr;  // device -> host copies of the body state run beside the next step
  cudaEvent_t evSnap = nullptr, evCopied = nullptr;
  bool copyPending = false;
  DevBuf<double> stD;  // snapshot of x R v w for the copy stream
  DevBuf<int> stI;     // sleeping, collection
  bool narrowTimed = false;  // events 16..19 were recorded by the last detect()
  bool evCreated = false;
  int coopBlocksV[2] = {0, 0};  // per kernel variant: [hub support]
  int coopBlocks = 0;     // co-resident CTAs for the cooperative PGS kernel (0: cooperative launch unsupported)
  int usePersistent = 1;  // 0 never, 1 heuristic, 2 always (AM3D_PGS_PERSISTENT)
  long long solveLaunches = 0;
  long long kernelLaunches = 0;
  double rowUpdates = 0, solveSeconds = 0;
};
