"""CPU tests: the oracle's self-consistency (the reference holds no golden vectors for the 3D path,
SURVEY.md §8c, so the 2D tests' patterns are re-expressed for 3D scenes as coarse checks)."""
import os

import numpy as np
import pytest

from tests.conftest import REF
from tests.util import mixed_scene, params, small_pile


def test_pile_settles_and_sleeps(oracle_lib):
    from oracle.oracle import Oracle
    o = Oracle(small_pile(), params(enable_merging=0))
    for _ in range(400):
        o.step(0.05)
    b = o.bodies()
    assert np.abs(b["v"]).max() < 1e-3
    assert b["x"][1:, 1].min() > 0.4  # nothing fell through the plane
    r = o.residuals()
    assert r[1] <= 1e-12  # friction box constraint holds exactly after the projection


def test_merging_collapses_pile_into_one_pinned_collection(oracle_lib):
    from oracle.oracle import Oracle
    o = Oracle(small_pile(), params())
    for _ in range(600):
        o.step(0.05)
    b = o.bodies()
    assert o.num_top_level() == 1
    assert (b["collection"] >= 0).all()
    ev = o.events()
    assert (ev[:, 1] == 0).sum() >= 36


def test_lcp_residuals_small_after_many_iterations(oracle_lib):
    from oracle.oracle import Oracle
    o = Oracle(small_pile(), params(enable_merging=0, iterations=2000, tolerance=1e-14))
    for _ in range(30):
        o.step(0.05)
    r = o.residuals()
    assert r[3] > 0
    assert r[0] < 1e-6 and r[1] <= 1e-12 and r[2] < 1e-5


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present")
def test_tower_xml_merges_with_plane(oracle_lib):
    from adaptivemerging_b200.scene import load_xml
    from oracle.oracle import Oracle
    o = Oracle(load_xml(os.path.join(REF, "scenes3D/tower.xml")))
    for _ in range(300):
        o.step(0.05)
    assert o.num_top_level() == 1
    assert o.bodies()["sleeping"].all()


def test_mixed_scene_runs(oracle_lib):
    from oracle.oracle import Oracle
    o = Oracle(mixed_scene(), params(enable_merging=0))
    seen = set()
    for _ in range(80):
        o.step(0.05)
        c = o.contacts()
        for k in range(len(c)):
            seen.add((int(c["bv1"][k]) >= 0, int(c["bv2"][k]) >= 0, int(c["bv2"][k]) == -2, int(c["leaf"][k]) >= 0))
    # tree x tree, tree x plane / box x tree and box pairs all occurred
    assert len(seen) >= 3
    assert np.isfinite(o.bodies()["x"]).all()


def test_config_d_dominos_cascade():
    """SURVEY.md 8d config D (paper Fig. 7): dominosPlatforms.xml, 600 steps to let the 99 dominos merge with their
    three sprung platforms, then the scripted push that replaces the README's mouse drag (omega = (0, 0, -2) on
    domino66, the leftmost domino of the top platform), 1400 more steps."""
    from adaptivemerging_b200.ctypes_defs import apply_overrides, default_params
    from oracle.oracle import Oracle
    from tests.util import golden_scene
    blob = golden_scene("dominosPlatforms")
    o = Oracle(blob, apply_overrides(default_params(), blob.overrides))
    for _ in range(600):
        o.step(0.05)
    ev = o.events()
    assert o.timings().n_bodies == 4 and o.timings().n_contacts == 0      # 3 platform collections + the plane
    assert (ev[:, 1] == 0).sum() == 99 and (ev[:, 1] == 1).sum() == 0     # every domino merged, nothing unmerged yet
    n0 = len(ev)
    o.add_body_velocity(blob.names.index("domino66"), domega=np.array([0.0, 0.0, -2.0]))
    for _ in range(100):
        o.step(0.05)
    ev = o.events()
    assert (ev[n0:, 1] == 1).sum() >= 60 and o.timings().n_bodies >= 60   # the push unmerges the top platform's row
    for _ in range(1300):
        o.step(0.05)
    ev = o.events()
    assert (ev[n0:, 1] == 0).sum() >= 300 and (ev[n0:, 1] == 1).sum() >= 300   # the cascade: unmerge / re-merge waves
    assert o.timings().n_bodies <= 6                                           # everything comes to rest merged again


def test_magnet_pulls_while_active(oracle_lib):
    """RigidBody.magnetic / activateMagnet (PGS.java:119,150,167): a box under a pinned magnetic slab hangs while the magnet
    is active (the multipliers of its contacts are not clamped) and falls when it is switched off (LCPApp3D key 7)."""
    from adaptivemerging_b200.ctypes_defs import default_params
    from adaptivemerging_b200.scene import SceneBuilder
    from oracle.oracle import Oracle
    sb = SceneBuilder()
    sb.add_plane((0, 0, 0), (0, 1, 0))
    slab = sb.add_box((4, 1, 4), (0, 3.0, 0), pinned=True, name="magnet")
    box = sb.add_box((1, 1, 1), (0.2, 2.01, -0.1), name="box")
    sb.bodies[slab].magnetic = True
    blob = sb.build()
    p = default_params()
    p.enable_merging = 0
    o = Oracle(blob, p)
    o.set_body_magnet(slab, 1)
    o.set_body_magnet(box, 1)      # not magnetic: ignored
    ys = []
    for s in range(110):
        if s == 60:
            o.set_body_magnet(slab, 0)
        o.step(0.05)
        ys.append(float(o.bodies()["x"][box, 1]))
    assert min(ys[:60]) > 1.9 and ys[-1] < 0.7
    # without the magnet the box drops at once
    o2 = Oracle(blob, p)
    for s in range(30):
        o2.step(0.05)
    assert o2.bodies()["x"][box, 1] < 1.5
