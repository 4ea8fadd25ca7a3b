/*
 * am3d.h — C ABI of the B200-native AdaptiveMerging 3D rigid-body step.
 *
 * This is the drop-in boundary for ONE path of EulalieCoevoet/AdaptiveMerging:
 * mergingBodies3D.RigidBodySystem.advanceTime(dt) and everything it calls
 * (reference: src/mergingBodies3D/RigidBodySystem.java:102-185).  The reference
 * has no FFI of its own; the seam is cut at the RigidBodySystem object as used
 * by LCPApp3D / XMLParser / Display.  The Java side keeps its
 * XML loader and UI and drives this library through Panama FFM or JNI; this
 * repo's tests drive it from Python ctypes.
 *
 * Conventions
 *  - plain C, no torch / CUDA types in any signature;
 *  - caller owns every host buffer, the library owns all device memory in ctx;
 *  - every entry point returns 0 (AM3D_OK) or a negative AM3D_E* code and never
 *    throws; am3d_last_error() gives a message valid until the next call;
 *  - matrices are 3x3 row-major doubles (m00 m01 m02 m10 ...), like
 *    javax.vecmath.Matrix3d field order;
 *  - a ctx is not re-entrant (the reference is single threaded); different
 *    ctxs are independent and may live on different GPUs.
 */
#ifndef AM3D_H
#define AM3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AM3D_OK 0
#define AM3D_EINVAL (-1)       /* bad argument / inconsistent scene            */
#define AM3D_ECUDA (-2)        /* CUDA runtime error (message in last_error)   */
#define AM3D_ENOGPU (-3)       /* no usable device: there is NO CPU fallback   */
#define AM3D_EUNSUPPORTED (-4) /* option of the reference not built yet        */
#define AM3D_ECAPACITY (-5)    /* a hard capacity limit was hit in the middle  \
                                  of a step (collection slots, colours, tree   \
                                  stack): the step is NOT rolled back, call    \
                                  am3d_reset before stepping again             */
#define AM3D_ESTATE (-6)       /* call order (e.g. step before upload_scene)   */

/* body types: XMLParser.parseBody (XMLParser.java:151-163) */
#define AM3D_BODY_BOX 0
#define AM3D_BODY_PLANE 1
#define AM3D_BODY_SPHERE 2
#define AM3D_BODY_MESH 3
#define AM3D_BODY_COMPOSITE 4

/* collision primitive types (what narrowPhase dispatches on,
 * CollisionProcessor.java:707-757) */
#define AM3D_SHAPE_BOX 0
#define AM3D_SHAPE_TREE 1
#define AM3D_SHAPE_PLANE 2

/* body flags */
#define AM3D_F_PINNED 1
#define AM3D_F_MAGNETIC 2
#define AM3D_F_MAGNET_ACTIVE 4
#define AM3D_F_SLEEPING 8
/* a body that is in the scene blob but not (yet) in RigidBodySystem.bodies: a pre-allocated clone of a factory part
 * (Factory.generateBody, Factory.java:99-116) waiting for am3d_activate_body, or a body taken out by am3d_remove_body */
#define AM3D_F_DORMANT 16

/* spring types: Spring.java:31 */
#define AM3D_SPRING_ZERO 0
#define AM3D_SPRING_WORLD 1
#define AM3D_SPRING_BODYBODY 2

/* contact states: Contact.java:56 */
#define AM3D_CS_BROKEN 0
#define AM3D_CS_ONEDGE 1
#define AM3D_CS_CLEAR 2

/* bv identity codes used in contact keys (Appendix B of SURVEY.md) */
#define AM3D_BV_NULL (-1)
#define AM3D_BV_PLANE_DUMMY (-2)

/*
 * The scene: what XMLParser.parse() (XMLParser.java:40-67) leaves in
 * RigidBodySystem.bodies / .springs, flattened to structure-of-arrays.
 * Bodies are in XML parse order (the reference's only body identity).
 * A body owns shape_count collision primitives starting at shape_first: one
 * for box / plane / sphere / mesh bodies, the parts (RigidBodyGeomComposite
 * .bodies, in order) for a composite.  Sphere trees are stored once per
 * distinct tree (instances of a mesh share it), breadth-first so that the
 * children of a node are contiguous; node_rank is the node's pre-order DFS
 * index inside its tree (the order the reference's recursion visits leaves).
 */
typedef struct am3d_scene {
  int32_t n_bodies;
  int32_t n_shapes;
  int32_t n_nodes;
  int32_t n_springs;
  int32_t n_scenes; /* batched independent copies; bodies of different scenes never interact */
  int32_t _pad0;

  /* per body [n_bodies] */
  const int32_t* body_type;
  const int32_t* body_flags;
  const int32_t* body_scene;       /* scene id in [0,n_scenes) */
  const int32_t* body_shape_first;
  const int32_t* body_shape_count;
  const double* body_x;            /* [3]  RigidBody.x   */
  const double* body_R;            /* [9]  RigidBody.theta */
  const double* body_v;            /* [3] */
  const double* body_omega;        /* [3] */
  const double* body_mass;         /* massLinear (kept non-zero for XML-pinned bodies, as in the reference) */
  const double* body_minv;
  const double* body_mass_angular0;/* [9] */
  const double* body_jinv0;        /* [9] */
  const double* body_friction;
  const double* body_restitution;
  const double* body_bbB;          /* [24] boundingBoxB, 8 points; unused for planes */
  const int32_t* body_bb_count;    /* 8, or 0 for planes (empty list in the reference) */

  /* per shape [n_shapes] */
  const int32_t* shape_type;
  const int32_t* shape_body;
  const double* shape_size;        /* [3] box: full side lengths (RigidBodyGeomBox.size); plane: n */
  const double* shape_radius;      /* box: |size/2| (XMLParser.java:392); plane: d; tree: unused */
  const double* shape_p;           /* [3] plane: p; else unused */
  const double* shape_B2C_R;       /* [9] part->composite transform (identity for simple bodies) */
  const double* shape_B2C_t;       /* [3] */
  const int32_t* shape_tree_root;  /* node index of the tree root, -1 if none */

  /* sphere-tree nodes [n_nodes] */
  const double* node_c;            /* [3] BVSphere.cB in the (part) body frame */
  const double* node_r;
  const int32_t* node_first_child; /* -1 for a leaf */
  const int32_t* node_child_count;
  const int32_t* node_rank;        /* pre-order DFS index within the tree */

  /* springs [n_springs]: Spring.java */
  const int32_t* spring_type;
  const int32_t* spring_body1;
  const int32_t* spring_body2;     /* -1 unless BODYBODY */
  const double* spring_pb1;        /* [3] */
  const double* spring_pb2;        /* [3] */
  const double* spring_pw;         /* [3] */
  const double* spring_k;
  const double* spring_d;
  const double* spring_l0;
  const double* spring_ls;
} am3d_scene;

/*
 * Every tunable of the path (SURVEY.md §5 "Config" / Appendix A), one POD.
 * am3d_default_params() fills the reference defaults.
 */
typedef struct am3d_params {
  /* CollisionProcessor.java:1058-1083, Contact.java:518 */
  int32_t warm_start;
  int32_t shuffle;                 /* CollisionProcessor.java:674: asks for an unspecified sweep order; accepted, the colour order is one */
  int32_t enable_post_stabilization; /* RigidBodySystem.java:354-377, PGS.java:86-89 */
  int32_t enable_compliance;
  int32_t collection_cd;           /* 0 brute force, 1 BVH, 2 sweep and prune (:768-790): pruning structures over the same member
                                      pair tests - identical contacts; detection here prunes by world AABB in every mode */
  int32_t restitution_override;
  int32_t friction_override;
  int32_t iterations;
  int32_t iterations_in_collection;
  int32_t _pad0;
  double feedback_stiffness;
  double compliance;
  double restitution;
  double friction;
  double tolerance;
  double omega;
  double sliding_threshold;
  /* RigidBodySystem.java:544-554 */
  int32_t use_gravity;
  int32_t use_coriolis;            /* RigidBodySystem.java:212-229, 295-304 */
  int32_t springs_enabled;         /* reference applies springs only if mouseSpring != null (:244-247) */
  int32_t _pad1;
  double gravity_amount;
  double gravity_angle_deg;
  double viscous_linear;
  double viscous_angular;
  double spring_k_mod;
  double spring_d_mod;
  /* Merging.java:30-54 */
  int32_t enable_merging;
  int32_t merge_pinned;
  int32_t merge_cycle_condition;   /* must be 0 (AM3D_EUNSUPPORTED): BodyPairContact.checkCyclesToUnmerge :375-394 recurses without end
                                      over a static set once a cycle has been broken - nothing defined to reproduce */
  int32_t merge_stable_contact;
  int32_t merge_let_it_breathe;
  int32_t enable_unmerging;
  int32_t unmerge_friction;
  int32_t unmerge_normal;
  int32_t unmerge_relative_motion;
  int32_t update_contacts_in_collections;
  int32_t organize_contacts;
  int32_t metric_position_level;   /* MotionMetricProcessor.java:75-116 */
  int32_t step_accum_merging;
  int32_t step_accum_unmerging;
  int32_t steps_between_merge;
  int32_t _pad2;
  double threshold_merge;
  double threshold_unmerge;
  double threshold_breath;
  /* Sleeping.java:25-30 */
  int32_t enable_sleeping;
  int32_t sleep_step_accum;
  double sleep_threshold;
} am3d_params;

/* The timing fields the reference's CSV / overlay read (RigidBodySystem.java:495-542),
 * measured with CUDA events per phase, in seconds, for the last step. */
typedef struct am3d_timings {
  int32_t n_bodies;      /* top-level bodies (a collection counts as one) */
  int32_t n_contacts;
  double detection;
  double warmstart;
  double lcp_solve;
  double update_collections;
  double contact_ordering;
  double single_it_pgs;
  double merging;
  double merging_build;
  double unmerging;
  double unmerging_build;
  double compute_time;
  /* extras (not in the CSV) */
  int32_t pgs_iterations; /* iterations actually executed by the full solve */
  int32_t pgs_colors;
  int32_t n_pairs;        /* broadphase candidate pairs */
  int32_t n_collections;
  double pgs_kernel_time; /* device time of the full-solve sweeps only */
  double narrowphase_kernel_time; /* device time of the narrowphase kernels (count + emit passes) */
  int32_t pgs_kernel;     /* sweep form of the last full solve: 0 k_pgs_color (one launch per phase), 1 k_pgs_persistent
                             (one cooperative launch), 2 k_pgs_cluster (one launch, a thread-block cluster per block of scenes) */
  int32_t pgs_giant_groups; /* groups of >= 65 contacts solved by k_pgs_giant in the last full solve */
} am3d_timings;

/* One contact as the tests and the Java mirror see it (Contact.java fields). */
typedef struct am3d_contact {
  int32_t body1, body2; /* leaf (or composite parent) body ids */
  int32_t csb1, csb2;   /* composite part index within the body, -1 = null */
  int32_t bv1, bv2;     /* global node index, AM3D_BV_NULL, AM3D_BV_PLANE_DUMMY */
  int32_t info;
  int32_t leaf;         /* box x tree: the tree leaf hit (not part of the identity) */
  int32_t state;        /* AM3D_CS_* */
  int32_t new_this_step;
  int32_t color;        /* colour of the full solve (-1 if not solved) */
  int32_t in_collection;/* 1 if this is an internal contact of a collection */
  int32_t hub_mask;     /* solve-order lists only: bit 0 / 1 = body1 / body2 side was treated as a hub */
  int32_t _pad;
  double contactB1[3], normalB1[3], tangent1B1[3], tangent2B1[3];
  double point_w[3], normal_w[3]; /* world frame at detection time */
  double violation, prev_violation;
  double lambda[3];
  double lambda_warm[3];
} am3d_contact;

/* One body pair (BodyPairContact.java) */
typedef struct am3d_bpc {
  int32_t body1, body2;
  int32_t in_collection;
  int32_t n_contacts;
  int32_t n_metric, n_state;
  double metric_hist[4];
  int32_t state_hist[4];
} am3d_bpc;

typedef struct am3d_ctx am3d_ctx;

/* lifecycle ------------------------------------------------------------- */
void am3d_default_params(am3d_params* p);
int am3d_create(int device, am3d_ctx** out);
int am3d_destroy(am3d_ctx* ctx);
const char* am3d_last_error(const am3d_ctx* ctx);
const char* am3d_version(void);

/* replaces XMLParser.parse() output hand-over + RigidBodySystem.reset() (:390-426) */
int am3d_upload_scene(am3d_ctx* ctx, const am3d_scene* scene);
int am3d_set_params(am3d_ctx* ctx, const am3d_params* p);
int am3d_get_params(const am3d_ctx* ctx, am3d_params* p);
int am3d_reset(am3d_ctx* ctx);

/* replaces RigidBodySystem.advanceTime(dt) (:102-185), nsteps times */
int am3d_step(am3d_ctx* ctx, double dt, int nsteps);
/* async variant for batched shards driven from one host thread: the steps are handed to the context's own worker
 * thread (which runs the host side of the step and its kernels) and the call returns at once; am3d_sync waits for
 * them and returns their status.  Any other call on the same ctx first waits for the pending steps. */
int am3d_step_async(am3d_ctx* ctx, double dt, int nsteps);
int am3d_sync(am3d_ctx* ctx);

/* UI hooks of the reference that write body state between steps
 * (LCPApp3D scripted pushes, MouseImpulse, Animation) */
int am3d_set_body_velocity(am3d_ctx* ctx, int body, const double v[3], const double omega[3]);
int am3d_add_body_velocity(am3d_ctx* ctx, int body, const double dv[3], const double domega[3]);
/* Animation.applyNonPersistant (Animation.java:82-160) writes `body.sleeping = false` next to the velocity it sets */
int am3d_set_body_sleeping(am3d_ctx* ctx, int body, int sleeping);
/* LCPApp3D.java:936-947 (key 7) toggles RigidBody.activateMagnet of the magnetic bodies (RigidBody.java:149-153,
 * XML tag <magnetic>, XMLParser.java:587); contacts of a body with an active magnet are solved without the
 * non-negativity / friction-cone clamps (PGS.java:119,150,167).  No effect on a body that is not magnetic. */
int am3d_set_body_magnet(am3d_ctx* ctx, int body, int active);
/* RigidBodySystem.add(body) (:78) as Factory.generateBody uses it (Factory.java:99-116): a DORMANT body of the scene blob
 * (a clone of a factory part) enters RigidBodySystem.bodies at the end of the list with the given state.
 * R = NULL: identity, v / omega = NULL: zero. */
int am3d_activate_body(am3d_ctx* ctx, int body, const double x[3], const double R[9], const double v[3], const double omega[3]);
/* RigidBodySystem.remove(body) (:383): the body leaves the simulation (it must not be part of a collection) */
int am3d_remove_body(am3d_ctx* ctx, int body);
/* MouseSpringForce (MouseSpringForce.java:69-101): a spring between the point grabPointB of `body` and the world point
 * pointW, applied in every applyExternalForces (RigidBodySystem.java:244-247) until released with body = -1.  The body
 * is woken and marked as picked (its body pairs lead the single sweep, CollisionProcessor.java:361-383). */
int am3d_set_mouse_spring(am3d_ctx* ctx, int body, const double grabPointB[3], const double pointW[3], double stiffness,
                          double damping, int apply_at_com);
/* MouseImpulse.apply (MouseImpulse.java:101-125) + Impulse (RigidBodySystem.java:249-267): at the next step a force
 * scale * |end - picked| along (picked - end) acts at pickedPointB on the body and its collection; the step after that
 * (or the re-application of the forces after a merge event) applies the stored force once more to the body alone, as
 * the reference does */
int am3d_apply_impulse(am3d_ctx* ctx, int body, const double pickedPointB[3], const double endPointW[3], double scale);
/* bulk version: per-body velocity increments [3n] each (zeros are skipped), added to the top-level entity.
 * The two arrays are uploaded on a separate stream and added by the NEXT am3d_step after its contact detection (which
 * reads positions only) - or by whichever other entry point is called first, so a read-back in between sees them applied.
 * Keep the host arrays unchanged until that call has returned (pinned memory: the copy is asynchronous). */
int am3d_add_velocities(am3d_ctx* ctx, const double* dv, const double* domega);
int am3d_upload_bodies(am3d_ctx* ctx, const double* x, const double* R, const double* v,
                       const double* omega); /* teacher forcing: overwrite the state of all leaf bodies */

/* outbound reads (Display.java:76-200, LCPApp3D.java:264-286) ---------------- */
int am3d_num_bodies(const am3d_ctx* ctx);
int am3d_download_bodies(am3d_ctx* ctx, double* x, double* R, double* v, double* omega,
                         int32_t* sleeping, int32_t* collection /* -1 or collection slot */);
/* Same read, not waited for: the state as of this call is snapshot on the device and copied to the (pinned) host
 * buffers on a second stream while the next am3d_step runs; the buffers are valid after am3d_wait_download() or the
 * next download call.  Lets a front end draw step N while step N+1 is computed. */
int am3d_download_bodies_async(am3d_ctx* ctx, double* x, double* R, double* v, double* omega,
                               int32_t* sleeping, int32_t* collection);
int am3d_wait_download(am3d_ctx* ctx);
int am3d_num_contacts(am3d_ctx* ctx, int include_internal);
int am3d_download_contacts(am3d_ctx* ctx, am3d_contact* out, int capacity, int include_internal,
                           int* count);
int am3d_num_bpcs(am3d_ctx* ctx);
int am3d_download_bpcs(am3d_ctx* ctx, am3d_bpc* out, int capacity, int* count);
int am3d_get_timings(am3d_ctx* ctx, am3d_timings* t);
int am3d_total_steps(const am3d_ctx* ctx);

/*
 * Phase-level entry points used by the parity tests (teacher forcing) and by
 * bench.py's roofline leg.  They run exactly the kernels am3d_step runs.
 */
/* broadphase + narrowphase + Contact.set from the current body state
 * (CollisionProcessor.collisionDetection :91-102) */
int am3d_detect(am3d_ctx* ctx);
/* the full solve (CollisionProcessor.solveLCP :108-137 → PGS.solve) on the
 * current contacts; fills lambda, deltaV and contact states */
int am3d_solve(am3d_ctx* ctx, double dt);
int am3d_download_deltav(am3d_ctx* ctx, double* dv /* [n_bodies*6] */);
/* overwrite the multipliers of the current external contacts (canonical order) before am3d_solve */
int am3d_set_lambdas(am3d_ctx* ctx, const double* lambda /* [count*3] */, int count);
/* counters: [0] kernels launched so far, [1] PGS sweep launches, [2] contact-row updates, [3] PGS kernel seconds */
int am3d_stats(am3d_ctx* ctx, double* out4);
/* device-side stopwatch on the context's stream: mark slot 0 (start) / 1 (stop), then read the span */
/* merge / unmerge decisions so far: rows (step, kind, bodyLo, bodyHi); kind 0 = the pair became internal to a
 * collection (Merging.merge, Merging.java:73-163), 1 = it left its collection (Merging.unmerge :215-374) */
int am3d_num_events(am3d_ctx* ctx);
int am3d_download_events(am3d_ctx* ctx, int32_t* out /* [capacity*4] */, int capacity, int* count);
/* tests: record the Gauss-Seidel sequence of every solve; which = 0 last full solve, 1 last single sweep, 2 last
 * post-stabilisation solve; out = NULL only reports the count */
int am3d_record_orders(am3d_ctx* ctx, int on);
int am3d_download_order(am3d_ctx* ctx, int which, am3d_contact* out, int capacity, int* count);
/* internal body pairs of collections (RigidCollection.bodyPairContacts with inCollection) */
int am3d_num_internal_bpcs(am3d_ctx* ctx);
int am3d_download_internal_bpcs(am3d_ctx* ctx, am3d_bpc* out, int capacity, int* count);
/* one RigidCollection: x[3] R[9] v[3] omega[3] mass minv jinv[9] massAngular[9] flags alive members stamp */
int am3d_download_collection(am3d_ctx* ctx, int slot, double* out42);
/* position of every leaf body's top-level entity (the body itself or its RigidCollection) in RigidBodySystem.bodies,
 * as a monotone key: sorting by it gives the list order the Java side has to mirror (Merging.java:105-110, :269-270) */
int am3d_download_list_order(am3d_ctx* ctx, int64_t* out /* [n_bodies] */);
/* engine options that are not reference parameters (none of them changes a result except hub_min_degree):
 *   "hub_min_degree"   body pairs per body from which a body is a hub of the contact graph, 0 = never; default 64
 *   "record_events"    merge / unmerge event log for am3d_download_events, default 1; 0 saves a read-back per merge step
 *   "pgs_persistent"   0 never / 1 heuristic / 2 always: the whole solve in one cooperative launch
 *   "pgs_clusters"     0/1: batched scenes partitioned over thread-block clusters (74 - 2 300 scenes)
 *   "pgs_tail_fusion"  0/1: trailing phases that hold one group per scene folded into one launch
 *   "pgs_fast_rows"    0/1: branch-free PGS row update (0 = the plain form everywhere)
 *   "giant_warps", "giant_chunk"   sphere-tree pairs with >= 65 contacts solved by a warp; contacts per chunk (512)
 *   "tree_split"       0/1: tree x tree pairs with a large frontier split into one narrowphase task per node pair
 *   "scene_bfs"        0/1: breadth-first layers of the single sweep per scene for batched contexts
 *   "own_primitives"   0/1: hand-written radix sort / prefix sum (1) or the CUB ones (0, the cross-check)
 *   "merge_exact_max_pairs"   mergeable pairs per component up to which Merging.merge is replayed in sequence (16 384) */
int am3d_set_option(am3d_ctx* ctx, const char* name, double value);
int am3d_mark(am3d_ctx* ctx, int slot);
int am3d_elapsed_ms(am3d_ctx* ctx, double* ms);
/* order (position in the Gauss-Seidel sequence) the full solve gave each
 * current external contact, for replaying on the CPU oracle */
int am3d_download_solve_order(am3d_ctx* ctx, int32_t* order, int capacity, int* count);

#ifdef __cplusplus
}
#endif
#endif /* AM3D_H */
