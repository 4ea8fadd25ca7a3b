"""Size-independent properties at sizes the oracle cannot reach (BASELINE.json's configs M and B): run-to-run
determinism (bit-identical states, contact identities and merge/unmerge events from two fresh contexts), the
upload -> download round trip, physical invariants of the 1M-box stack, and agreement of the batched copies."""
import numpy as np
import pytest

from adaptivemerging_b200.ctypes_defs import apply_overrides, contact_keys, default_params
from adaptivemerging_b200.scene import box_stack
from adaptivemerging_b200.system import RigidBodySystem
from tests.util import golden_scene, params

pytestmark = pytest.mark.gpu


def _run(blob, p, steps):
    s = RigidBodySystem(0).load(blob, p)
    s.advanceTime(0.05, steps)
    b = s.bodies()
    c = s.contacts()
    ev = s.events().tolist()
    t = s.timings()
    s.close()
    return b, c, ev, t


@pytest.mark.parametrize("workload", ["stack32_merging", "batch32"])
def test_run_to_run_determinism(workload):
    if workload == "stack32_merging":
        blob, p, steps = box_stack(32, 32, 32, pile=True), params(), 25
    else:
        one = golden_scene("tower25platform")
        blob, p, steps = one.replicate(32), apply_overrides(default_params(), one.overrides), 170
    b1, c1, e1, t1 = _run(blob, p, steps)
    b2, c2, e2, t2 = _run(blob, p, steps)
    for k in ("x", "R", "v", "omega"):
        assert np.array_equal(b1[k].view(np.uint64), b2[k].view(np.uint64)), k
    assert np.array_equal(b1["collection"], b2["collection"]) and np.array_equal(b1["sleeping"], b2["sleeping"])
    assert np.array_equal(contact_keys(c1), contact_keys(c2))
    assert np.array_equal(c1["lambda"].view(np.uint64), c2["lambda"].view(np.uint64))
    assert e1 == e2
    assert t1.n_contacts == t2.n_contacts > 0


def test_body_state_round_trip():
    blob = box_stack(20, 20, 20, pile=True)
    s = RigidBodySystem(0).load(blob, params(enable_merging=0))
    s.advanceTime(0.05, 5)
    b = s.bodies()
    s.upload_bodies(b["x"], b["R"], b["v"], b["omega"])
    b2 = s.bodies()
    for k in ("x", "R", "v", "omega"):
        assert np.array_equal(b[k].view(np.uint64), b2[k].view(np.uint64)), k


def test_million_box_stack_invariants():
    """config M: 100 x 100 columns x 100 layers, merging off, a dozen steps."""
    n = 100
    blob = box_stack(n, n, n)
    s = RigidBodySystem(0).load(blob, params(enable_merging=0))
    s.advanceTime(0.05, 12)
    b = s.bodies()
    t = s.timings()
    box = blob.a["body_type"] != 1  # everything but the plane
    x = b["x"][box]
    assert np.isfinite(b["x"]).all() and np.isfinite(b["v"]).all() and np.isfinite(b["R"]).all()
    assert t.n_contacts >= 4 * n * n * n * 0.95          # >= one face (4 contacts) per box: a resting stack
    assert t.pgs_iterations == 30
    # nothing fell through the plane, nothing was ejected, columns did not drift
    assert x[:, 1].min() > 0.45 and x[:, 1].max() < n + 1.0
    x0 = blob.a["body_x"].reshape(-1, 3)[box]
    assert np.abs(x[:, [0, 2]] - x0[:, [0, 2]]).max() < 0.05
    # rotations stay orthonormal (normalizeCP) and velocities small
    R = b["R"][box].reshape(-1, 3, 3)
    assert np.abs(np.einsum("nij,nkj->nik", R, R) - np.eye(3)).max() < 1e-9
    assert np.abs(b["v"][box]).max() < 2.0
    c = s.contacts()
    assert c["violation"].min() > -0.2
    s.close()


def test_batched_copies_agree():
    """config B: every copy is its own scene, coloured and solved (one CTA per scene, own tolerance exit) exactly as if it
    were alone: identical copies stay bit-identical, and they never interact."""
    one = golden_scene("tower25platform")
    copies = 64
    blob = one.replicate(copies)
    p = apply_overrides(default_params(), one.overrides)
    s = RigidBodySystem(0).load(blob, p)
    s.advanceTime(0.05, 140)
    b = s.bodies()
    nb = one.a["body_type"].shape[0]
    for k in ("x", "R", "v", "omega"):
        a = b[k].reshape(copies, nb, -1)
        assert np.array_equal(a.view(np.uint64), np.broadcast_to(a[:1], a.shape).view(np.uint64)), k
    c = s.contacts()
    assert (c["body1"] // nb == c["body2"] // nb).all()  # no contact crosses scenes
    counts = np.bincount(c["body1"] // nb, minlength=copies)
    assert counts.min() == counts.max() > 0
    s.close()


@pytest.mark.parametrize("copies", [8, 80])
def test_scene_in_a_batch_equals_the_scene_alone(copies):
    """Scene k of a batched context == the same scene run alone in its own context, bit for bit: body states, contact
    multipliers, merge / unmerge events and PGS iteration counts.  The copies get different initial kicks, so they
    collapse differently and leave the PGS at different iteration counts (per-scene tolerance exit, PGS.java:190-192).
    8 copies run the grid-wide sweeps, 80 copies the partitioned ones (one thread-block cluster per block of scenes)."""
    one = golden_scene("tower25platform")
    steps = 170
    nb = one.a["body_type"].shape[0]
    p = apply_overrides(default_params(), one.overrides)
    kicked = one.names.index("B3L0")

    def kick(blob, scene, k):
        v = blob.a["body_v"].reshape(-1, 3)
        v[scene * nb + kicked] = (0.03 * k, 0.0, -0.02 * k)

    batch = one.replicate(copies)
    for k in range(copies):
        kick(batch, k, k % 8)
    s = RigidBodySystem(0).load(batch, p)
    s.advanceTime(0.05, steps)
    bb, cb, eb = s.bodies(), s.contacts(), s.events()
    s_launches = s.stats()["solve_launches"]
    s.close()
    assert len(eb) > 0
    for k in (0, 3, copies - 1):
        alone = one.replicate(1)
        kick(alone, 0, k % 8)
        a = RigidBodySystem(0).load(alone, p)
        a.advanceTime(0.05, steps)
        ba, ca, ea = a.bodies(), a.contacts(), a.events()
        a.close()
        sl = slice(k * nb, (k + 1) * nb)
        for f in ("x", "R", "v", "omega"):
            assert np.array_equal(bb[f][sl].view(np.uint64), ba[f].view(np.uint64)), (k, f)
        assert np.array_equal(bb["sleeping"][sl], ba["sleeping"])
        mine = cb[(cb["body1"] // nb) == k]
        assert len(mine) == len(ca)
        assert np.array_equal(mine["lambda"].view(np.uint64), ca["lambda"].view(np.uint64)), k
        ev = eb[(eb[:, 2] // nb) == k].copy()
        ev[:, 2:] -= k * nb
        assert sorted(map(tuple, ev.tolist())) == sorted(map(tuple, ea.tolist())), k
    # the copies really differ
    x = bb["x"].reshape(copies, nb, 3)
    assert np.abs(x[7] - x[0]).max() > 1e-3
    if copies >= 80:
        assert s_launches < 0.5 * steps * 60, s_launches  # one launch per solve, not one per phase
