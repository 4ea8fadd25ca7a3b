"""The reference's off-by-default options, each in lockstep with the oracle's restatement of it."""
import numpy as np
import pytest

from adaptivemerging_b200.ctypes_defs import apply_overrides
from tests.util import golden_scene, lockstep, mixed_scene, params, same_partition, small_pile

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scene", ["pile", "mixed", "tower25platform"])
def test_post_stabilization_lockstep(scene):
    """enablePostStabilization (RigidBodySystem.java:354-377, PGS.java:86-89, CollisionProcessor.java:119): second detection
    at the advanced positions, position-level solve, bodies moved by deltaV; the velocity solve loses its Baumgarte term."""
    blob = {"pile": small_pile, "mixed": mixed_scene, "tower25platform": lambda: golden_scene("tower25platform")}[scene]()
    p = apply_overrides(params(), blob.overrides)
    p.enable_post_stabilization = 1
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, p, 160 if scene != "tower25platform" else 130, tol=1e-6)
    assert ev_g == ev_o
    assert same_partition(gpu.bodies()["collection"], cpu.bodies()["collection"])
    assert gpu.max_contacts > 0
    # it does what it is for: penetration stays small without the velocity-level feedback
    c = gpu.contacts()
    if len(c) and scene != "tower25platform":  # (the towers are collapsing at that point)
        assert c["violation"].min() > -0.05


@pytest.mark.parametrize("scene", ["mixed", "tower"])
def test_coriolis_lockstep(scene):
    """useCoriolis (RigidBodySystem.java:212-229 gyroscopic stabilisation of massAngular, :295-304 Coriolis torque on every
    unpinned body, members of collections included; both applied again when a merge event re-applies the forces)."""
    blob = mixed_scene() if scene == "mixed" else golden_scene("tower")
    p = apply_overrides(params(), blob.overrides)
    p.use_coriolis = 1
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, p, 200, tol=1e-6)
    assert ev_g == ev_o
    assert same_partition(gpu.bodies()["collection"], cpu.bodies()["collection"])
    # and it is not a no-op: the same run without the term ends elsewhere
    p0 = apply_overrides(params(), blob.overrides)
    from adaptivemerging_b200.system import RigidBodySystem
    ref = RigidBodySystem(0).load(blob, p0)
    ref.advanceTime(0.05, 200)
    assert np.abs(ref.bodies()["omega"] - gpu.bodies()["omega"]).max() > 0 or np.abs(ref.bodies()["x"] - gpu.bodies()["x"]).max() > 0


@pytest.mark.parametrize("scene", ["pile", "tower25platform"])
def test_position_level_metric_lockstep(scene):
    """metricPositionLevel (MotionMetricProcessor.java:75-116, BodyPairContact.java:92-93, :134-135): the merge / unmerge motion
    metric is the displacement of the bounding-box points in the other body's frame over one position update"""
    blob = small_pile() if scene == "pile" else golden_scene("tower25platform")
    p = apply_overrides(params(), blob.overrides)
    p.metric_position_level = 1
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, p, 200 if scene == "pile" else 150, tol=1e-6)
    assert ev_g == ev_o and len(ev_g) > 0
    assert same_partition(gpu.bodies()["collection"], cpu.bodies()["collection"])


def test_unorganized_sweep_lockstep():
    """organizeContacts = false (CollisionProcessor.java:259-272): the single sweep takes the external contacts, then the
    internal contacts of the awake collections, instead of the breadth-first order"""
    from tests.util import pile_with_bullet
    p = params()
    p.organize_contacts = 0
    gpu, cpu, ev_g, ev_o, worst = lockstep(pile_with_bullet(), p, 420, tol=1e-5)
    assert ev_g == ev_o
    assert any(e[1] == 1 for e in ev_g)


def test_shuffle_is_accepted():
    """shuffle asks for an unspecified sweep order (Collections.shuffle, unseeded): the colour order is one, deterministically"""
    from adaptivemerging_b200.system import RigidBodySystem
    p = params()
    p.shuffle = 1
    a = RigidBodySystem(0).load(small_pile(), p)
    b = RigidBodySystem(0).load(small_pile(), params())
    a.advanceTime(0.05, 60); b.advanceTime(0.05, 60)
    assert np.array_equal(a.bodies()["x"], b.bodies()["x"])
