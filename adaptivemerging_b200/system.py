"""Host-side mirror of the reference's ``RigidBodySystem`` for the hot path.

Same names, argument meaning and error behaviour as ``mergingBodies3D.RigidBodySystem``
(RigidBodySystem.java:25-66, 102-185, 390-446): ``advanceTime(dt)``, ``reset()``, ``clear()``,
``jiggle()`` and the timing fields the overlay / CSV read (``computeTime``, ``mergingTime``,
``unmergingTime``, ``warmStartTime``, ``totalSteps``; ``collision.collisionDetectTime`` ... are exposed
through ``timings()``).  The reference is Java; with no JVM in this environment the host side above the
C ABI is written in Python and calls the CUDA library through ctypes exactly as the Java shim in
INTEGRATION.md would through Panama FFM.  All physics runs on the GPU; this class holds no state beyond
the opaque context handle.
"""
import ctypes as C

import numpy as np

from . import _capi
from .ctypes_defs import (BPC_DTYPE, CONTACT_DTYPE, am3d_params, am3d_timings, apply_overrides, default_params)
from .csvlog import CsvLog
from .scene import SceneBlob, load_xml


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RigidBodySystem:
    def __init__(self, device=0):
        self._L = _capi.load()
        h = C.c_void_p()
        rc = self._L.am3d_create(int(device), C.byref(h))
        if rc != 0:
            raise _capi.Am3dError(rc, "am3d_create failed (no CUDA device? there is no CPU fallback)")
        self._h = h
        self.device = device
        self.params = default_params()
        self.blob = None
        self.name = ""
        self.simulationTime = 0.0
        self.totalSteps = 0
        self.computeTime = 0.0
        self.totalAccumulatedComputeTime = 0.0
        self.mergingTime = 0.0
        self.unmergingTime = 0.0
        self.warmStartTime = 0.0
        self.saveCSV = False        # RigidBodySystem.saveCSV (:547): one CSV row per advanceTime, exportDataToFile :495-542
        self.sceneName = "scene"
        self._csv = CsvLog()

    # -- plumbing ---------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            msg = self._L.am3d_last_error(self._h)
            raise _capi.Am3dError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_csv", None):
            self._csv.close()
        if getattr(self, "_h", None):
            self._L.am3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene hand-over (what XMLParser.parse leaves in the system) ---------------------------------
    def load(self, blob: SceneBlob, params: am3d_params = None):
        self.blob = blob
        self._scene = blob.as_ctypes()
        self.params = params if params is not None else apply_overrides(default_params(), blob.overrides)
        self._ck(self._L.am3d_set_params(self._h, C.byref(self.params)))
        self._ck(self._L.am3d_upload_scene(self._h, C.byref(self._scene)))
        self.simulationTime = 0.0
        self.totalSteps = 0
        return self

    def loadXML(self, path, copies=1):
        return self.load(load_xml(path, copies=copies))

    def set_params(self, params: am3d_params):
        self._ck(self._L.am3d_set_params(self._h, C.byref(params)))
        self.params = params

    # -- RigidBodySystem API -----------------------------------------------------------------------
    def advanceTime(self, dt, nsteps=1):
        self._ck(self._L.am3d_step(self._h, float(dt), int(nsteps)))
        t = self.timings()
        self.totalSteps += nsteps
        self.simulationTime += dt * nsteps
        self.computeTime = t.compute_time
        self.totalAccumulatedComputeTime += t.compute_time
        self.mergingTime = t.merging
        self.unmergingTime = t.unmerging
        self.warmStartTime = t.warmstart
        if self.saveCSV or self._csv.stream is not None:
            self._csv.export(self.saveCSV, self.sceneName, bool(self.params.enable_merging), t)

    def step_async(self, dt, nsteps=1):
        """hand nsteps to the context's worker thread and return (am3d_step_async); sync() waits"""
        self._ck(self._L.am3d_step_async(self._h, float(dt), int(nsteps)))
        self.totalSteps += nsteps
        self.simulationTime += dt * nsteps

    def sync(self):
        self._ck(self._L.am3d_sync(self._h))

    def reset(self):
        self._ck(self._L.am3d_reset(self._h))
        self.simulationTime = 0.0
        self.totalAccumulatedComputeTime = 0.0
        self.totalSteps = 0

    def clear(self):
        """RigidBodySystem.clear (:441-446): drop the scene."""
        self.close()
        self.__init__(self.device)

    def jiggle(self, seed=None):
        """RigidBodySystem.jiggle (:85-96): random velocity kick on every unpinned body."""
        rng = np.random.default_rng(seed)
        b = self.bodies()
        pinned = (self.blob.a["body_flags"] & 1) != 0
        kick_w = rng.random((self.n_bodies, 3)) * 2 - 1
        kick_v = rng.random((self.n_bodies, 3)) * 2 - 1
        kick_w[pinned] = 0
        kick_v[pinned] = 0
        self.upload_bodies(b["x"], b["R"], b["v"] + kick_v, b["omega"] + kick_w)

    # -- outbound reads ----------------------------------------------------------------------------
    @property
    def n_bodies(self):
        return self._L.am3d_num_bodies(self._h)

    def bodies(self, out=None):
        """Body state of the leaf bodies.  `out`: a dict returned by an earlier call (or caller-owned, e.g. pinned,
        arrays of the same shapes) to download into without allocating."""
        n = self.n_bodies
        if out is not None:
            x, R, v, w, sl, co = out["x"], out["R"], out["v"], out["omega"], out["sleeping"], out["collection"]
            self._ck(self._L.am3d_download_bodies(self._h, _p(x), _p(R), _p(v), _p(w), _p(sl), _p(co)))
            return out
        x = np.empty((n, 3)); R = np.empty((n, 9)); v = np.empty((n, 3)); w = np.empty((n, 3))
        sl = np.empty(n, np.int32); co = np.empty(n, np.int32)
        self._ck(self._L.am3d_download_bodies(self._h, _p(x), _p(R), _p(v), _p(w), _p(sl), _p(co)))
        return dict(x=x, R=R, v=v, omega=w, sleeping=sl, collection=co)

    def bodies_async(self, out):
        """Start the download of the body state into `out` (pinned arrays as returned by bodies()); it completes under
        the next advanceTime.  wait_bodies() (or the next download) makes `out` valid."""
        x, R, v, w, sl, co = out["x"], out["R"], out["v"], out["omega"], out["sleeping"], out["collection"]
        self._ck(self._L.am3d_download_bodies_async(self._h, _p(x), _p(R), _p(v), _p(w), _p(sl), _p(co)))
        return out

    def wait_bodies(self):
        self._ck(self._L.am3d_wait_download(self._h))

    def upload_bodies(self, x, R, v, omega):
        x, R, v, omega = [np.ascontiguousarray(a, np.float64) for a in (x, R, v, omega)]
        self._ck(self._L.am3d_upload_bodies(self._h, _p(x), _p(R), _p(v), _p(omega)))

    def set_body_velocity(self, body, v=None, omega=None):
        v = np.ascontiguousarray(v, np.float64) if v is not None else None
        w = np.ascontiguousarray(omega, np.float64) if omega is not None else None
        self._ck(self._L.am3d_set_body_velocity(self._h, int(body), _p(v), _p(w)))

    def add_body_velocity(self, body, dv=None, domega=None):
        v = np.ascontiguousarray(dv, np.float64) if dv is not None else None
        w = np.ascontiguousarray(domega, np.float64) if domega is not None else None
        self._ck(self._L.am3d_add_body_velocity(self._h, int(body), _p(v), _p(w)))

    # -- Factory / UI hooks (RigidBodySystem.add / remove, MouseSpringForce, MouseImpulse, Animation) ----------------
    def set_body_sleeping(self, body, sleeping):
        self._ck(self._L.am3d_set_body_sleeping(self._h, int(body), int(bool(sleeping))))

    def set_body_magnet(self, body, active):
        """RigidBody.activateMagnet as LCPApp3D's key 7 toggles it (:936-947)"""
        self._ck(self._L.am3d_set_body_magnet(self._h, int(body), int(bool(active))))

    def activate_body(self, body, x, R=None, v=None, omega=None):
        """RigidBodySystem.add of a dormant clone (Factory.generateBody)"""
        a = [np.ascontiguousarray(q, np.float64) if q is not None else None for q in (x, R, v, omega)]
        self._ck(self._L.am3d_activate_body(self._h, int(body), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3])))

    def remove_body(self, body):
        self._ck(self._L.am3d_remove_body(self._h, int(body)))

    def set_mouse_spring(self, body, grab_point_b=None, point_w=None, stiffness=50.0, damping=10.0, at_com=False):
        g = np.ascontiguousarray(grab_point_b, np.float64) if grab_point_b is not None else None
        w = np.ascontiguousarray(point_w, np.float64) if point_w is not None else None
        self._ck(self._L.am3d_set_mouse_spring(self._h, -1 if body is None else int(body), _p(g), _p(w), float(stiffness), float(damping), int(at_com)))

    def apply_impulse(self, body, picked_point_b, end_point_w, scale=1.0):
        g = np.ascontiguousarray(picked_point_b, np.float64)
        w = np.ascontiguousarray(end_point_w, np.float64)
        self._ck(self._L.am3d_apply_impulse(self._h, int(body), _p(g), _p(w), float(scale)))

    def add_velocities(self, dv, domega):
        """bulk velocity pokes, [n,3] each (the batched form of MouseImpulse / scripted pushes)"""
        dv = np.ascontiguousarray(dv, np.float64)
        dw = np.ascontiguousarray(domega, np.float64)
        self._ck(self._L.am3d_add_velocities(self._h, _p(dv), _p(dw)))

    def contacts(self, include_internal=False):
        n = self._L.am3d_num_contacts(self._h, int(include_internal))
        out = np.zeros(max(n, 1), CONTACT_DTYPE)
        cnt = C.c_int(0)
        self._ck(self._L.am3d_download_contacts(self._h, _p(out), n, int(include_internal), C.byref(cnt)))
        return out[:cnt.value]

    def bpcs(self):
        n = self._L.am3d_num_bpcs(self._h)
        out = np.zeros(max(n, 1), BPC_DTYPE)
        cnt = C.c_int(0)
        self._ck(self._L.am3d_download_bpcs(self._h, _p(out), n, C.byref(cnt)))
        return out[:cnt.value]

    def timings(self) -> am3d_timings:
        t = am3d_timings()
        self._ck(self._L.am3d_get_timings(self._h, C.byref(t)))
        return t

    def stats(self):
        out = np.zeros(4)
        self._ck(self._L.am3d_stats(self._h, _p(out)))
        return dict(kernel_launches=int(out[0]), solve_launches=int(out[1]), row_updates=out[2], solve_seconds=out[3])

    def events(self):
        """Merge / unmerge decisions so far: rows (step, kind, bodyLo, bodyHi); kind 0 = pair became internal to a
        collection (Merging.merge), 1 = pair left its collection (Merging.unmerge)."""
        n = self._L.am3d_num_events(self._h)
        out = np.zeros((max(n, 1), 4), np.int32)
        cnt = C.c_int(0)
        self._ck(self._L.am3d_download_events(self._h, _p(out), n, C.byref(cnt)))
        return out[:cnt.value]

    def record_orders(self, on=True):
        self._ck(self._L.am3d_record_orders(self._h, int(on)))

    def order(self, which):
        """Gauss-Seidel sequence (contact identities) of the last full solve (0) / single sweep (1) / post-stabilisation solve (2)."""
        cnt = C.c_int(0)
        self._ck(self._L.am3d_download_order(self._h, int(which), None, 0, C.byref(cnt)))  # size query
        cap = cnt.value + 1
        out = np.zeros(cap, CONTACT_DTYPE)
        self._ck(self._L.am3d_download_order(self._h, int(which), _p(out), cap, C.byref(cnt)))
        return out[:cnt.value]

    def internal_bpcs(self):
        n = self._L.am3d_num_internal_bpcs(self._h)
        out = np.zeros(max(n, 1), BPC_DTYPE)
        cnt = C.c_int(0)
        self._ck(self._L.am3d_download_internal_bpcs(self._h, _p(out), n, C.byref(cnt)))
        return out[:cnt.value]

    def collection(self, slot):
        out = np.zeros(42)
        self._ck(self._L.am3d_download_collection(self._h, int(slot), _p(out)))
        return dict(x=out[0:3], R=out[3:12], v=out[12:15], omega=out[15:18], mass=out[18], minv=out[19], jinv=out[20:29],
                    mA=out[29:38], flags=int(out[38]), alive=int(out[39]), count=int(out[40]), stamp=int(out[41]))

    def list_order(self, raw=False):
        """rank of every leaf body's top-level entity in RigidBodySystem.bodies (0 = first in the list); raw: the monotone
        position keys themselves (dormant bodies carry a stale key)"""
        out = np.zeros(self.n_bodies, np.int64)
        self._ck(self._L.am3d_download_list_order(self._h, _p(out)))
        return out if raw else np.unique(out, return_inverse=True)[1]

    def set_option(self, name, value):
        self._ck(self._L.am3d_set_option(self._h, name.encode(), float(value)))

    def mark(self, slot):
        self._ck(self._L.am3d_mark(self._h, int(slot)))

    def elapsed_ms(self):
        ms = C.c_double(0)
        self._ck(self._L.am3d_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    # -- phase-level entry points (parity tests, roofline leg of bench.py) ------------------------------
    def detect(self):
        self._ck(self._L.am3d_detect(self._h))
        return self._L.am3d_num_contacts(self._h, 0)

    def set_lambdas(self, lam):
        lam = np.ascontiguousarray(lam, np.float64)
        self._ck(self._L.am3d_set_lambdas(self._h, _p(lam), int(lam.size // 3)))

    def solve(self, dt=0.05):
        self._ck(self._L.am3d_solve(self._h, float(dt)))

    def deltav(self):
        dv = np.empty((self.n_bodies, 6))
        self._ck(self._L.am3d_download_deltav(self._h, _p(dv)))
        return dv

    def solve_order(self):
        n = self._L.am3d_num_contacts(self._h, 0)
        out = np.zeros(max(n, 1), np.int32)
        cnt = C.c_int(0)
        self._ck(self._L.am3d_download_solve_order(self._h, _p(out), n, C.byref(cnt)))
        return out[:cnt.value]
