"""GPU parity with adaptive merging on: merge / unmerge decisions and trajectories against the oracle.

The oracle replays the GPU's Gauss-Seidel sequences (full solve and single sweep).  Decisions must be
IDENTICAL over the stated horizon; body states agree to 1e-6 (collection mass properties are summed in a
different floating-point order on the GPU, SURVEY.md "Hard parts" 7, and sin/cos differ by ulps).
"""
import numpy as np
import pytest

from tests.util import mixed_scene, params, small_pile

pytestmark = pytest.mark.gpu


from tests.util import lockstep, same_partition  # noqa: E402


def test_pile_merges_identically():
    gpu, cpu, ev_g, ev_o, worst = lockstep(small_pile(), params(), 320)
    assert ev_g == ev_o
    assert len(ev_g) > 30
    assert same_partition(gpu.bodies()["collection"], cpu.bodies()["collection"])


def test_merge_then_unmerge_cascade():
    # the pile merges with the plane and falls asleep; a box dropped from high above then hits it: the
    # collection must wake and split the same way on both sides
    from tests.util import pile_with_bullet
    gpu, cpu, ev_g, ev_o, worst = lockstep(pile_with_bullet(), params(), 420, tol=1e-5)
    assert ev_g == ev_o
    assert any(e[1] == 1 for e in ev_g), "no unmerge happened"
    assert same_partition(gpu.bodies()["collection"], cpu.bodies()["collection"])


def test_mixed_scene_with_merging():
    gpu, cpu, ev_g, ev_o, worst = lockstep(mixed_scene(), params(), 200)
    assert ev_g == ev_o


@pytest.mark.parametrize("scene", ["plate", "tower25platform"])
def test_hub_mode_lockstep(scene):
    """bodies touched by >= 8 body pairs (the plate / the sprung platform) are solved as hubs of the contact graph:
    merging decisions must still match the oracle, which replays the hub sequence"""
    from tests.util import golden_scene, hub_scene
    from adaptivemerging_b200.ctypes_defs import apply_overrides
    blob = hub_scene() if scene == "plate" else golden_scene("tower25platform")
    p = apply_overrides(params(), blob.overrides)
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, p, 150, tol=1e-6, options={"hub_min_degree": 8})
    assert gpu.hub_contacts > 0
    assert ev_g == ev_o


def test_config_d_dominos_cascade_lockstep():
    """600 steps of dominosPlatforms.xml (everything merges with the sprung platforms), the
    scripted push on domino66, 1400 more steps of the unmerge / re-merge cascade; identical events on both sides.  The
    single sweep runs in the reference's breadth-first order (getOrganizedContacts, CollisionProcessor.java:346-441),
    permuted only inside a layer; the oracle checks that (tests/util.py::lockstep, oracle updateInCollections)."""
    from adaptivemerging_b200.ctypes_defs import apply_overrides, default_params
    from tests.util import golden_scene
    blob = golden_scene("dominosPlatforms")
    p = apply_overrides(default_params(), blob.overrides)
    poke = {600: (blob.names.index("domino66"), None, np.array([0.0, 0.0, -2.0]))}
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, p, 2000, poke=poke, tol=1e-5)
    assert ev_g == ev_o
    assert sum(1 for e in ev_g if e[1] == 1) >= 60
    assert same_partition(gpu.bodies()["collection"], cpu.bodies()["collection"])
