"""The hand-written radix sort / prefix sum (csrc/am3d_sort.cuh) against the CUB primitives they replace: the same scenes
stepped both ways give the same run bit for bit (broadphase cell and pair sorts, solve order, warm-start indices, merge /
unmerge bookkeeping all go through them)."""
import numpy as np
import pytest

from adaptivemerging_b200.ctypes_defs import apply_overrides, default_params
from adaptivemerging_b200.scene import box_stack
from adaptivemerging_b200.system import RigidBodySystem
from tests.util import golden_scene, params

pytestmark = pytest.mark.gpu


def run(blob, p, steps, own):
    s = RigidBodySystem(0).load(blob, p)
    s.set_option("own_primitives", own)
    s.advanceTime(0.05, steps)
    out = (s.bodies(), s.events().tolist(), s.timings().n_contacts, s.contacts(True))
    s.close()
    return out


@pytest.mark.parametrize("name,steps", [("tower25platform", 170), ("torsos", 125), ("dominosPlatforms", 150)])
def test_reference_scenes_same_run_with_own_and_library_primitives(name, steps):
    blob = golden_scene(name)
    p = apply_overrides(default_params(), blob.overrides)
    a = run(blob, p, steps, 1)
    b = run(blob, p, steps, 0)
    assert a[1] == b[1] and a[2] == b[2] and a[2] > 0
    for k in ("x", "R", "v", "omega", "sleeping", "collection"):
        assert np.array_equal(a[0][k], b[0][k]), k
    assert a[3].tobytes() == b[3].tobytes()


def test_batched_pile_same_run_with_own_and_library_primitives():
    """48 scenes in one context: partitioned sweeps, tail phases, scene-keyed cell codes; sizes that are not multiples of a tile"""
    blob = box_stack(3, 5, 3, pile=True).replicate(48)
    a = run(blob, params(), 90, 1)
    b = run(blob, params(), 90, 0)
    assert a[1] == b[1] and a[2] == b[2] and len(a[1]) > 0
    for k in ("x", "R", "v", "omega"):
        assert np.array_equal(a[0][k], b[0][k]), k
