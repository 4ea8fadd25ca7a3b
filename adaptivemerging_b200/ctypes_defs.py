"""ctypes mirrors of the PODs declared in include/am3d.h (am3d_params, am3d_timings, am3d_contact,
am3d_bpc).  Field order must match the header exactly."""
import ctypes as C


class am3d_params(C.Structure):
    _fields_ = [
        ("warm_start", C.c_int32), ("shuffle", C.c_int32), ("enable_post_stabilization", C.c_int32),
        ("enable_compliance", C.c_int32), ("collection_cd", C.c_int32), ("restitution_override", C.c_int32),
        ("friction_override", C.c_int32), ("iterations", C.c_int32), ("iterations_in_collection", C.c_int32),
        ("_pad0", C.c_int32),
        ("feedback_stiffness", C.c_double), ("compliance", C.c_double), ("restitution", C.c_double),
        ("friction", C.c_double), ("tolerance", C.c_double), ("omega", C.c_double), ("sliding_threshold", C.c_double),
        ("use_gravity", C.c_int32), ("use_coriolis", C.c_int32), ("springs_enabled", C.c_int32), ("_pad1", C.c_int32),
        ("gravity_amount", C.c_double), ("gravity_angle_deg", C.c_double), ("viscous_linear", C.c_double),
        ("viscous_angular", C.c_double), ("spring_k_mod", C.c_double), ("spring_d_mod", C.c_double),
        ("enable_merging", C.c_int32), ("merge_pinned", C.c_int32), ("merge_cycle_condition", C.c_int32),
        ("merge_stable_contact", C.c_int32), ("merge_let_it_breathe", C.c_int32), ("enable_unmerging", C.c_int32),
        ("unmerge_friction", C.c_int32), ("unmerge_normal", C.c_int32), ("unmerge_relative_motion", C.c_int32),
        ("update_contacts_in_collections", C.c_int32), ("organize_contacts", C.c_int32),
        ("metric_position_level", C.c_int32), ("step_accum_merging", C.c_int32), ("step_accum_unmerging", C.c_int32),
        ("steps_between_merge", C.c_int32), ("_pad2", C.c_int32),
        ("threshold_merge", C.c_double), ("threshold_unmerge", C.c_double), ("threshold_breath", C.c_double),
        ("enable_sleeping", C.c_int32), ("sleep_step_accum", C.c_int32), ("sleep_threshold", C.c_double),
    ]


def default_params() -> am3d_params:
    """Reference defaults: CollisionProcessor.java:1058-1083, Contact.java:518,
    RigidBodySystem.java:544-554, Merging.java:30-54, Sleeping.java:25-30."""
    p = am3d_params()
    p.warm_start = 1
    p.enable_compliance = 1
    p.iterations = 30
    p.iterations_in_collection = 1
    p.feedback_stiffness = 0.5
    p.compliance = 1e-3
    p.restitution = 0.5
    p.friction = 0.1
    p.tolerance = 1e-5
    p.omega = 1.0
    p.sliding_threshold = 0.01
    p.use_gravity = 1
    p.springs_enabled = 1
    p.gravity_amount = 1.0
    p.gravity_angle_deg = 90.0
    p.viscous_linear = 1.0
    p.viscous_angular = 1.0
    p.spring_k_mod = 1.0
    p.spring_d_mod = 1.0
    p.enable_merging = 1
    p.merge_pinned = 1
    p.merge_stable_contact = 1
    p.merge_let_it_breathe = 1
    p.enable_unmerging = 1
    p.unmerge_friction = 1
    p.unmerge_normal = 1
    p.unmerge_relative_motion = 1
    p.update_contacts_in_collections = 1
    p.organize_contacts = 1
    p.step_accum_merging = 3
    p.step_accum_unmerging = 3
    p.steps_between_merge = 10
    p.threshold_merge = 1e-2
    p.threshold_unmerge = 2e-2
    p.threshold_breath = 1e-5
    p.enable_sleeping = 1
    p.sleep_step_accum = 10
    p.sleep_threshold = 1e-5
    return p


def apply_overrides(p: am3d_params, overrides: dict) -> am3d_params:
    """XML <collision>/<system> attribute overrides (XMLParser.java:94-134)."""
    for k, v in overrides.items():
        if hasattr(p, k):
            setattr(p, k, v)
    return p


class am3d_timings(C.Structure):
    _fields_ = [
        ("n_bodies", C.c_int32), ("n_contacts", C.c_int32),
        ("detection", C.c_double), ("warmstart", C.c_double), ("lcp_solve", C.c_double),
        ("update_collections", C.c_double), ("contact_ordering", C.c_double), ("single_it_pgs", C.c_double),
        ("merging", C.c_double), ("merging_build", C.c_double), ("unmerging", C.c_double),
        ("unmerging_build", C.c_double), ("compute_time", C.c_double),
        ("pgs_iterations", C.c_int32), ("pgs_colors", C.c_int32), ("n_pairs", C.c_int32), ("n_collections", C.c_int32),
        ("pgs_kernel_time", C.c_double), ("narrowphase_kernel_time", C.c_double),
        ("pgs_kernel", C.c_int32), ("pgs_giant_groups", C.c_int32),
    ]


class am3d_contact(C.Structure):
    _fields_ = [
        ("body1", C.c_int32), ("body2", C.c_int32), ("csb1", C.c_int32), ("csb2", C.c_int32),
        ("bv1", C.c_int32), ("bv2", C.c_int32), ("info", C.c_int32), ("leaf", C.c_int32),
        ("state", C.c_int32), ("new_this_step", C.c_int32), ("color", C.c_int32), ("in_collection", C.c_int32),
        ("hub_mask", C.c_int32), ("_pad", C.c_int32),
        ("contactB1", C.c_double * 3), ("normalB1", C.c_double * 3), ("tangent1B1", C.c_double * 3),
        ("tangent2B1", C.c_double * 3), ("point_w", C.c_double * 3), ("normal_w", C.c_double * 3),
        ("violation", C.c_double), ("prev_violation", C.c_double),
        ("lambda_", C.c_double * 3), ("lambda_warm", C.c_double * 3),
    ]


class am3d_bpc(C.Structure):
    _fields_ = [
        ("body1", C.c_int32), ("body2", C.c_int32), ("in_collection", C.c_int32), ("n_contacts", C.c_int32),
        ("n_metric", C.c_int32), ("n_state", C.c_int32),
        ("metric_hist", C.c_double * 4), ("state_hist", C.c_int32 * 4),
    ]


import numpy as np  # noqa: E402

CONTACT_DTYPE = np.dtype([
    ("body1", "<i4"), ("body2", "<i4"), ("csb1", "<i4"), ("csb2", "<i4"), ("bv1", "<i4"), ("bv2", "<i4"),
    ("info", "<i4"), ("leaf", "<i4"), ("state", "<i4"), ("new_this_step", "<i4"), ("color", "<i4"),
    ("in_collection", "<i4"), ("hub_mask", "<i4"), ("_pad", "<i4"),
    ("contactB1", "<f8", 3), ("normalB1", "<f8", 3), ("tangent1B1", "<f8", 3), ("tangent2B1", "<f8", 3),
    ("point_w", "<f8", 3), ("normal_w", "<f8", 3), ("violation", "<f8"), ("prev_violation", "<f8"),
    ("lambda", "<f8", 3), ("lambda_warm", "<f8", 3),
])
BPC_DTYPE = np.dtype([
    ("body1", "<i4"), ("body2", "<i4"), ("in_collection", "<i4"), ("n_contacts", "<i4"), ("n_metric", "<i4"),
    ("n_state", "<i4"), ("metric_hist", "<f8", 4), ("state_hist", "<i4", 4),
])
assert CONTACT_DTYPE.itemsize == C.sizeof(am3d_contact), (CONTACT_DTYPE.itemsize, C.sizeof(am3d_contact))
assert BPC_DTYPE.itemsize == C.sizeof(am3d_bpc), (BPC_DTYPE.itemsize, C.sizeof(am3d_bpc))

KEY_FIELDS = ["body1", "body2", "csb1", "csb2", "bv1", "bv2", "info", "leaf"]


def contact_keys(arr):
    """(n,8) int array of the full identity of each contact (warm-start key + box×tree leaf)."""
    return np.stack([arr[f] for f in KEY_FIELDS], axis=1)
