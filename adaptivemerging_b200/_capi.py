"""ctypes binding of the product library ``libam3d.so`` (C ABI declared in include/am3d.h).

There is no CPU fallback: if the CUDA library is missing or no GPU is present the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AM3D_LIBRARY") or os.path.join(_HERE, "libam3d.so")  # (AM3D_LIBRARY: another build of the same library, for A/B runs)

OK, EINVAL, ECUDA, ENOGPU, EUNSUPPORTED, ECAPACITY, ESTATE = 0, -1, -2, -3, -4, -5, -6

EXPORTS = [
    "am3d_default_params", "am3d_create", "am3d_destroy", "am3d_last_error", "am3d_version", "am3d_upload_scene",
    "am3d_set_params", "am3d_get_params", "am3d_reset", "am3d_step", "am3d_step_async", "am3d_sync",
    "am3d_set_body_velocity", "am3d_add_body_velocity", "am3d_upload_bodies", "am3d_num_bodies",
    "am3d_download_bodies", "am3d_num_contacts", "am3d_download_contacts", "am3d_num_bpcs", "am3d_download_bpcs",
    "am3d_get_timings", "am3d_total_steps", "am3d_detect", "am3d_solve",
    "am3d_download_deltav", "am3d_set_lambdas", "am3d_stats", "am3d_download_solve_order", "am3d_mark", "am3d_elapsed_ms",
    "am3d_num_events", "am3d_download_events", "am3d_record_orders", "am3d_download_order", "am3d_num_internal_bpcs",
    "am3d_download_internal_bpcs", "am3d_download_collection", "am3d_set_option", "am3d_add_velocities",
    "am3d_download_bodies_async", "am3d_wait_download", "am3d_download_list_order", "am3d_set_body_sleeping",
    "am3d_activate_body", "am3d_remove_body", "am3d_set_mouse_spring", "am3d_apply_impulse", "am3d_set_body_magnet",
]

_LIB = None


class Am3dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"am3d error {code}: {msg}")
        self.code = code


def load():
    """Load libam3d.so; raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C adaptivemerging_b200/csrc` "
                              "(there is no CPU fallback for the rigid-body step)")
        L = C.CDLL(LIB_PATH)
        L.am3d_last_error.restype = C.c_char_p
        L.am3d_last_error.argtypes = [C.c_void_p]
        L.am3d_version.restype = C.c_char_p
        L.am3d_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.am3d_step.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.am3d_step_async.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.am3d_solve.argtypes = [C.c_void_p, C.c_double]
        L.am3d_mark.argtypes = [C.c_void_p, C.c_int]
        L.am3d_record_orders.argtypes = [C.c_void_p, C.c_int]
        L.am3d_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.am3d_download_collection.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.am3d_num_contacts.argtypes = [C.c_void_p, C.c_int]
        L.am3d_set_body_sleeping.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.am3d_set_body_magnet.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.am3d_activate_body.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.am3d_remove_body.argtypes = [C.c_void_p, C.c_int]
        L.am3d_set_mouse_spring.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int]
        L.am3d_apply_impulse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double]
        for n in ["am3d_destroy", "am3d_sync", "am3d_reset", "am3d_num_bodies", "am3d_num_bpcs", "am3d_total_steps",
                  "am3d_detect", "am3d_num_events", "am3d_num_internal_bpcs"]:
            getattr(L, n).argtypes = [C.c_void_p]
        _LIB = L
    return _LIB
