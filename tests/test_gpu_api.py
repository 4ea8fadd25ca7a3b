"""The C ABI's contract (include/am3d.h): error codes instead of exceptions, call-order checks, unsupported reference
options refused loudly, and reset() = RigidBodySystem.reset (:390-410): the run after a reset repeats the first
run bit for bit."""
import numpy as np
import pytest

from adaptivemerging_b200 import _capi
from adaptivemerging_b200.ctypes_defs import apply_overrides, default_params
from adaptivemerging_b200.system import RigidBodySystem
from tests.util import golden_scene, params, small_pile

pytestmark = pytest.mark.gpu


def _code(exc):
    return exc.value.code


def test_call_order_and_bad_arguments():
    s = RigidBodySystem(0)
    with pytest.raises(_capi.Am3dError) as e:
        s.advanceTime(0.05)
    assert _code(e) == _capi.ESTATE           # step before upload_scene
    s.load(small_pile(), params())
    with pytest.raises(_capi.Am3dError) as e:
        s.set_body_velocity(10 ** 6, v=np.zeros(3))
    assert _code(e) == _capi.EINVAL
    s.advanceTime(0.05, 3)
    assert s.timings().n_contacts > 0
    s.close()


def test_unsupported_reference_options_are_refused():
    s = RigidBodySystem(0)
    p = default_params()
    p.merge_cycle_condition = 1
    with pytest.raises(_capi.Am3dError) as e:
        s.set_params(p)
    assert _code(e) == _capi.EUNSUPPORTED
    p = default_params()
    p.collection_cd = 3   # CollisionProcessor.java:788 "Unsupported collision detection method"
    with pytest.raises(_capi.Am3dError) as e:
        s.set_params(p)
    assert _code(e) == _capi.EINVAL
    s.close()


def test_collection_collision_modes_give_the_same_run():
    """collectionCD 1 (BVH) and 2 (sweep and prune) prune the member-pair tests of mode 0 with enclosing volumes
    (CollisionProcessor.java:768-790, 859-960): same contacts, same run."""
    blob = small_pile(4, 5, 4)
    runs = []
    for mode in (0, 1, 2):
        p = params()
        p.collection_cd = mode
        s = RigidBodySystem(0).load(blob, p)
        s.advanceTime(0.05, 120)
        runs.append((s.bodies(), s.events().tolist(), s.timings().n_contacts))
        s.close()
    assert len(runs[0][1]) > 0
    for b, ev, nc in runs[1:]:
        assert ev == runs[0][1] and nc == runs[0][2]
        for k in ("x", "R", "v", "omega"):
            assert np.array_equal(b[k], runs[0][0][k]), k


def test_reset_repeats_the_run_bit_for_bit():
    blob = small_pile(4, 5, 4)   # merges into one pinned collection with the plane within ~40 steps
    p = params()
    s = RigidBodySystem(0).load(blob, p)
    s.advanceTime(0.05, 120)
    first, ev1 = s.bodies(), s.events().tolist()
    s.reset()
    assert s.totalSteps == 0
    b0 = s.bodies()
    assert np.array_equal(b0["x"], blob.a["body_x"].reshape(-1, 3)) and (b0["collection"] < 0).all()
    s.advanceTime(0.05, 120)
    second, ev2 = s.bodies(), s.events().tolist()
    for k in ("x", "R", "v", "omega"):
        assert np.array_equal(first[k].view(np.uint64), second[k].view(np.uint64)), k
    assert np.array_equal(first["collection"], second["collection"])
    assert ev1 == ev2 and len(ev1) > 0
    s.close()


def test_step_async_returns_before_the_steps_are_done():
    """am3d_step_async hands the steps to the context's worker thread: the call returns at once, am3d_sync waits, and the
    result is the one am3d_step gives; two contexts driven from one host thread overlap."""
    import time
    blob = golden_scene("tower25platform")
    p = apply_overrides(default_params(), blob.overrides)
    a = RigidBodySystem(0).load(blob, p)
    b = RigidBodySystem(0).load(blob, p)
    a.advanceTime(0.05, 5); b.advanceTime(0.05, 5)      # warm both contexts
    t0 = time.perf_counter(); a.advanceTime(0.05, 150); t_sync = time.perf_counter() - t0
    t0 = time.perf_counter()
    b.step_async(0.05, 150)
    t_call = time.perf_counter() - t0
    b.sync()
    t_total = time.perf_counter() - t0
    assert t_call < 0.2 * t_sync, (t_call, t_sync)       # returned long before 150 steps could have run
    assert t_total > 0.5 * t_sync
    ga, gb = a.bodies(), b.bodies()
    for k in ("x", "R", "v", "omega"):
        assert np.array_equal(ga[k].view(np.uint64), gb[k].view(np.uint64)), k
    assert a.events().tolist() == b.events().tolist()
    a.close(); b.close()
