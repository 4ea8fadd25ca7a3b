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


def _run_forms(blob, p, steps, options):
    s = RigidBodySystem(0).load(blob, p)
    for k, v in options.items():
        s.set_option(k, v)
    s.advanceTime(0.05, steps)
    out = (s.bodies(), s.events().tolist(), s.timings().n_contacts, s.contacts(True), s.stats()["solve_launches"])
    s.close()
    return out


def test_sweep_forms_agree_bit_for_bit_on_batched_towers():
    """80 copies of tower25platform.xml (the platform's 13 pairs force trailing phases with one group per scene): the
    launch-per-phase sweep with the folded tail (k_pgs_tail), the same without folding, the cluster form and the
    cooperative form walk the same Gauss-Seidel sequence - bodies, events and contact records (multipliers included)
    are identical; so are the branch-free and the plain row update, and the per-scene and the grid layering of the sweep."""
    blob = golden_scene("tower25platform").replicate(80)   # enough scenes for the cluster form (>= 2 x resident clusters) and the per-scene layering (>= 32)
    p = apply_overrides(default_params(), blob.overrides)
    steps = 150   # collections form around step 130: the single sweep takes part
    phase = {"pgs_persistent": 0, "pgs_clusters": 0}
    ref = _run_forms(blob, p, steps, dict(phase, pgs_tail_fusion=0))
    assert len(ref[1]) > 0 and ref[2] > 0
    forms = {"tail": dict(phase, pgs_tail_fusion=1), "default": {}, "persistent": {"pgs_persistent": 2, "pgs_clusters": 0},
             "plain rows": dict(phase, pgs_fast_rows=0), "grid layering": {"scene_bfs": 0}}
    for name, opt in forms.items():
        r = _run_forms(blob, p, steps, opt)
        assert r[1] == ref[1] and r[2] == ref[2], name
        for k in ("x", "R", "v", "omega", "sleeping", "collection"):
            assert np.array_equal(r[0][k], ref[0][k]), (name, k)
        assert r[3]["lambda"].tobytes() == ref[3]["lambda"].tobytes(), name
        if name == "tail":
            assert r[4] < ref[4], "the tail phases were not folded"   # fewer sweep launches per step
