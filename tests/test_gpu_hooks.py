"""The callers' hooks into the step: Factory.generateBody / RigidBodySystem.add and remove, MouseSpringForce,
MouseImpulse, Animation's sleeping flag - each in lockstep with the oracle's restatement."""
import numpy as np
import pytest

from adaptivemerging_b200 import _capi
from adaptivemerging_b200.scene import _boxes_blob, box_stack
from adaptivemerging_b200.system import RigidBodySystem
from tests.util import lockstep, params, same_partition

pytestmark = pytest.mark.gpu

DORMANT = 16


def pile_with_dormant_clones(n_clones=6):
    """a 2x3x2 pile plus n dormant unit boxes parked far away (the pre-allocated clones of a factory part)"""
    base = box_stack(2, 3, 2, pile=True)
    xs = base.a["body_x"].reshape(-1, 3)[1:].copy()
    Rs = base.a["body_R"].reshape(-1, 3, 3)[1:].copy()
    n0 = len(xs)
    xs = np.concatenate([xs, np.tile([[500.0, 500.0, 500.0]], (n_clones, 1))])
    Rs = np.concatenate([Rs, np.tile(np.eye(3)[None], (n_clones, 1, 1))])
    blob = _boxes_blob(np.ones((len(xs), 3)), xs, Rs)
    first = 1 + n0   # body 0 is the plane
    blob.a["body_flags"][first:] |= DORMANT
    return blob, first


def test_factory_generates_bodies():
    """Factory.generateBody (Factory.java:99-116): every 12 steps a dormant clone enters the simulation above the pile with a
    pose and velocity chosen by the caller (the Java side keeps the java.util.Random draws)."""
    blob, first = pile_with_dormant_clones(6)
    rng = np.random.default_rng(3)
    script = {}
    for k in range(6):
        x = np.array([rng.uniform(-0.6, 0.6), 7.0, rng.uniform(-0.6, 0.6)])
        a = rng.uniform(0, 1)
        R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
        v = np.array([rng.uniform(-0.1, 0.1), -1.0, rng.uniform(-0.1, 0.1)])
        w = rng.uniform(-0.2, 0.2, 3)
        script[20 + 12 * k] = (lambda s, b=first + k, x=x, R=R, v=v, w=w: s.activate_body(b, x, R, v, w))
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, params(), 260, tol=1e-6, script=script)
    assert ev_g == ev_o and len(ev_g) > 0
    assert same_partition(gpu.bodies()["collection"], cpu.bodies()["collection"])
    y = gpu.bodies()["x"][first:, 1]
    assert (y < 6.0).all() and (y > 0.4).all()          # they fell onto the pile and stayed above the plane
    with pytest.raises(_capi.Am3dError):                 # already in the simulation
        gpu.activate_body(first, np.zeros(3))


def test_remove_body():
    blob, first = pile_with_dormant_clones(1)
    script = {5: lambda s: s.activate_body(first, np.array([3.0, 0.6, 3.0])), 40: lambda s: s.remove_body(first)}
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, params(), 80, script=script)
    assert ev_g == ev_o
    c = gpu.contacts()
    assert not ((c["body1"] == first) | (c["body2"] == first)).any()


def test_mouse_spring_pulls_a_merged_body_out():
    """MouseSpringForce.apply (:69-101): the pile merges and falls asleep, then the mouse spring grabs the top box: it is woken,
    marked as picked (its pairs lead the single sweep, CollisionProcessor.java:361-383), pulled, and unmerges."""
    blob = box_stack(2, 3, 2, pile=True)
    top = int(np.argmax(blob.a["body_x"].reshape(-1, 3)[:, 1]))
    grab = np.array([0.3, 0.2, -0.1])
    script = {150: lambda s: s.set_mouse_spring(top, grab, np.array([1.5, 6.0, 0.5]), 50.0, 10.0, False),
              175: lambda s: s.set_mouse_spring(top, grab, np.array([-2.0, 5.0, 1.0]), 50.0, 10.0, True),
              200: lambda s: s.set_mouse_spring(None)}
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, params(), 260, tol=1e-5, script=script)
    assert ev_g == ev_o
    assert any(e[1] == 1 and e[0] > 150 for e in ev_g), "the grabbed box never left its collection"


def test_mouse_impulse():
    """MouseImpulse.apply (:101-125) + the stored Impulse applied once more at the following applyExternalForces"""
    blob = box_stack(2, 3, 2, pile=True)
    top = int(np.argmax(blob.a["body_x"].reshape(-1, 3)[:, 1]))
    script = {120: lambda s: s.apply_impulse(top, np.array([0.1, 0.4, 0.0]), np.array([3.0, 2.5, 0.0]), 4.0)}
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, params(), 200, tol=1e-5, script=script)
    assert ev_g == ev_o
    assert np.abs(gpu.bodies()["x"][top] - blob.a["body_x"].reshape(-1, 3)[top]).max() > 0.05


def test_animation_style_velocity_writes():
    """Animation.applyNonPersistant (Animation.java:82-160): velocity components written every step + sleeping = false"""
    blob = box_stack(2, 2, 2, pile=True)
    b = 3
    script = {}
    for k in range(90, 110):
        script[k] = (lambda s, r=(k - 89) / 20.0: (s.set_body_sleeping(b, 0), s.set_body_velocity(b, np.array([0.6 * r, 0.0, 0.0]), None)))
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, params(enable_merging=0), 140, script=script)
    assert ev_g == ev_o


def test_magnet_holds_a_box_until_it_is_switched_off():
    """RigidBody.magnetic / activateMagnet (RigidBody.java:149-153; toggled by LCPApp3D's key 7, :936-947): the contacts of a
    body with an active magnet are solved without the clamps (PGS.java:119,150,167), so the multiplier can pull."""
    from adaptivemerging_b200.scene import F_MAGNETIC, SceneBuilder
    sb = SceneBuilder()
    sb.add_plane((0, 0, 0), (0, 1, 0))
    slab = sb.add_box((4, 1, 4), (0, 3.0, 0), pinned=True, name="magnet")
    box = sb.add_box((1, 1, 1), (0.2, 2.01, -0.1), name="box")     # its top face sits 0.01 inside the slab's bottom face
    sb.bodies[slab].magnetic = True
    blob = sb.build()
    assert blob.a["body_flags"][slab] & F_MAGNETIC
    script = {0: lambda s: s.set_body_magnet(slab, 1), 60: lambda s: s.set_body_magnet(slab, 0),
              3: lambda s: s.set_body_magnet(box, 1)}                # not magnetic: ignored
    heights = []
    p = params(enable_merging=0)

    class Probe:   # record the height on the GPU side every step through the script hook
        def __call__(self, s):
            if hasattr(s, "_h"):
                heights.append(float(s.bodies()["x"][box, 1]))
    probe = Probe()
    full = {k: (lambda s, f=script.get(k): (probe(s), f(s) if f else None)) for k in range(120)}
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, p, 120, tol=1e-6, script=full)
    assert ev_g == ev_o
    assert min(heights[:60]) > 1.9, "the active magnet did not hold the box"
    assert heights[-1] < 0.7, "the box did not fall after the magnet was switched off"
    # and with merging on: same events on both sides
    gpu, cpu, ev_g, ev_o, worst = lockstep(blob, params(), 120, tol=1e-6, script=script)
    assert ev_g == ev_o


def test_velocity_pokes_uploaded_beside_detection_equal_per_body_writes():
    """am3d_add_velocities: the per-body pokes go up on the copy stream and are applied by the next step after its
    detection (or by the next call, whichever comes first) - the run equals one with the same pokes written body by body."""
    blob = box_stack(3, 4, 3, pile=True)
    nb = blob.n_bodies
    rng = np.random.default_rng(11)
    a = RigidBodySystem(0).load(blob, params())
    b = RigidBodySystem(0).load(blob, params())
    for step in range(60):
        if step % 7 == 3:
            dv = np.zeros((nb, 3)); dw = np.zeros((nb, 3))
            for body in rng.choice(np.arange(1, nb), 3, replace=False):
                dv[body] = rng.uniform(-0.3, 0.3, 3); dw[body] = rng.uniform(-0.2, 0.2, 3)
            a.add_velocities(dv, dw)
            if step == 10:   # read back before stepping: the pending poke is applied first
                va = a.bodies()["v"].copy()
            for body in np.nonzero(np.abs(dv).sum(1) + np.abs(dw).sum(1))[0]:
                b.add_body_velocity(int(body), dv[body], dw[body])
            if step == 10:
                assert np.array_equal(va, b.bodies()["v"])
        a.advanceTime(0.05)
        b.advanceTime(0.05)
    ga, gb = a.bodies(), b.bodies()
    for k in ("x", "R", "v", "omega"):
        assert np.array_equal(ga[k], gb[k]), k
    assert a.events().tolist() == b.events().tolist()
    a.close(); b.close()
