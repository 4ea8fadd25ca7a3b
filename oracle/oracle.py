"""ctypes binding of oracle/libam_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (adaptivemerging_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from adaptivemerging_b200.ctypes_defs import (BPC_DTYPE, CONTACT_DTYPE, am3d_bpc, am3d_contact, am3d_params,
                                              am3d_timings, apply_overrides, default_params)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libam_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.amo_create.restype = C.c_void_p
        L.amo_create.argtypes = [C.c_void_p, C.c_void_p]
        for name in ["amo_destroy", "amo_set_params", "amo_get_bodies", "amo_set_bodies", "amo_get_timings",
                     "amo_apply_external_forces", "amo_set_lambdas", "amo_get_deltav", "amo_set_next_orders",
                     "amo_get_events", "amo_set_body_velocity", "amo_add_body_velocity", "amo_residuals", "amo_get_list_order", "amo_set_next_order_post",
                     "amo_set_body_sleeping", "amo_set_body_magnet", "amo_activate_body", "amo_remove_body", "amo_set_mouse_spring", "amo_apply_impulse"]:
            getattr(L, name).restype = None
        for name in ["amo_row_updates", "amo_solve_seconds"]:
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p]
        L.amo_step.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.amo_set_mouse_spring.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int]
        L.amo_apply_impulse.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double]
        L.amo_activate_body.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.amo_solve.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """The CPU restatement of RigidBodySystem for one scene blob."""

    def __init__(self, blob, params=None):
        self.L = lib()
        self.blob = blob
        self.params = params if params is not None else apply_overrides(default_params(), blob.overrides)
        self._scene = blob.as_ctypes()
        self.h = C.c_void_p(self.L.amo_create(C.byref(self._scene), C.byref(self.params)))
        self.n = blob.n_bodies

    def __del__(self):
        try:
            if self.h:
                self.L.amo_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_params(self, p):
        self.params = p
        self.L.amo_set_params(self.h, C.byref(p))

    def step(self, dt=0.05, n=1):
        return self.L.amo_step(self.h, dt, n)

    def bodies(self):
        n = self.n
        x = np.empty((n, 3)); R = np.empty((n, 9)); v = np.empty((n, 3)); w = np.empty((n, 3))
        sl = np.empty(n, np.int32); co = np.empty(n, np.int32)
        self.L.amo_get_bodies(self.h, _p(x), _p(R), _p(v), _p(w), _p(sl), _p(co))
        return dict(x=x, R=R, v=v, omega=w, sleeping=sl, collection=co)

    def set_bodies(self, x, R, v, omega):
        x, R, v, omega = [np.ascontiguousarray(a, np.float64) for a in (x, R, v, omega)]
        self.L.amo_set_bodies(self.h, _p(x), _p(R), _p(v), _p(omega))

    def contacts(self, include_internal=False):
        n = self.L.amo_num_contacts(self.h, int(include_internal))
        out = np.zeros(max(n, 1), CONTACT_DTYPE)
        m = self.L.amo_get_contacts(self.h, _p(out), n, int(include_internal))
        return out[:m]

    def bpcs(self, include_internal=False):
        cap = self.L.amo_num_bpcs(self.h) + (1 << 16 if include_internal else 0)
        out = np.zeros(max(cap, 1), BPC_DTYPE)
        m = self.L.amo_get_bpcs(self.h, _p(out), cap, int(include_internal))
        return out[:m]

    def timings(self):
        t = am3d_timings()
        self.L.amo_get_timings(self.h, C.byref(t))
        return t

    def detect(self):
        return self.L.amo_detect(self.h)

    def apply_external_forces(self):
        self.L.amo_apply_external_forces(self.h)

    def solve(self, dt=0.05, order=None):
        if order is None:
            return self.L.amo_solve(self.h, dt, None, 0)
        order = np.ascontiguousarray(order, CONTACT_DTYPE)
        return self.L.amo_solve(self.h, dt, _p(order), len(order))

    def set_lambdas(self, lam):
        lam = np.ascontiguousarray(lam, np.float64)
        self.L.amo_set_lambdas(self.h, _p(lam))

    def deltav(self):
        dv = np.empty((self.n, 6))
        self.L.amo_get_deltav(self.h, _p(dv))
        return dv

    def set_next_orders(self, full=None, sweep=None, post=None):
        f = np.ascontiguousarray(full, CONTACT_DTYPE) if full is not None else None
        s = np.ascontiguousarray(sweep, CONTACT_DTYPE) if sweep is not None else None
        q = np.ascontiguousarray(post, CONTACT_DTYPE) if post is not None else None
        self.L.amo_set_next_order_post(self.h, _p(q), len(q) if q is not None else 0)
        self.L.amo_set_next_orders(self.h, _p(f), len(f) if f is not None else 0, _p(s), len(s) if s is not None else 0)

    def events(self):
        n = self.L.amo_num_events(self.h)
        out = np.zeros((max(n, 1), 4), np.int32)
        self.L.amo_get_events(self.h, _p(out))
        return out[:n]

    def set_body_velocity(self, body, v=None, omega=None):
        v = np.ascontiguousarray(v, np.float64) if v is not None else None
        w = np.ascontiguousarray(omega, np.float64) if omega is not None else None
        self.L.amo_set_body_velocity(self.h, body, _p(v), _p(w))

    def add_body_velocity(self, body, dv=None, domega=None):
        v = np.ascontiguousarray(dv, np.float64) if dv is not None else None
        w = np.ascontiguousarray(domega, np.float64) if domega is not None else None
        self.L.amo_add_body_velocity(self.h, body, _p(v), _p(w))

    def set_body_sleeping(self, body, sleeping):
        self.L.amo_set_body_sleeping(self.h, int(body), int(bool(sleeping)))

    def set_body_magnet(self, body, active):
        self.L.amo_set_body_magnet(self.h, int(body), int(bool(active)))

    def activate_body(self, body, x, R=None, v=None, omega=None):
        a = [np.ascontiguousarray(q, np.float64) if q is not None else None for q in (x, R, v, omega)]
        self.L.amo_activate_body(self.h, int(body), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]))

    def remove_body(self, body):
        self.L.amo_remove_body(self.h, int(body))

    def set_mouse_spring(self, body, grab_point_b=None, point_w=None, stiffness=50.0, damping=10.0, at_com=False):
        g = np.ascontiguousarray(grab_point_b, np.float64) if grab_point_b is not None else None
        w = np.ascontiguousarray(point_w, np.float64) if point_w is not None else None
        self.L.amo_set_mouse_spring(self.h, -1 if body is None else int(body), _p(g), _p(w), float(stiffness), float(damping), int(at_com))

    def apply_impulse(self, body, picked_point_b, end_point_w, scale=1.0):
        g = np.ascontiguousarray(picked_point_b, np.float64)
        w = np.ascontiguousarray(end_point_w, np.float64)
        self.L.amo_apply_impulse(self.h, int(body), _p(g), _p(w), float(scale))

    def list_order(self):
        out = np.zeros(self.n, np.int32)
        self.L.amo_get_list_order(self.h, _p(out))
        return out

    def residuals(self):
        out = np.zeros(4)
        self.L.amo_residuals(self.h, _p(out))
        return out

    def num_top_level(self):
        return self.L.amo_num_top_level(self.h)

    def row_updates(self):
        return self.L.amo_row_updates(self.h)

    def solve_seconds(self):
        return self.L.amo_solve_seconds(self.h)
