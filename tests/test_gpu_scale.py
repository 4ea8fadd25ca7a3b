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
    """config B: every copy is its own scene; copies only differ in Gauss-Seidel order (colour priorities hash the
    group index), so before the towers collapse they agree to solver precision and they never interact."""
    one = golden_scene("tower25platform")
    copies = 64
    blob = one.replicate(copies)
    p = apply_overrides(default_params(), one.overrides)
    s = RigidBodySystem(0).load(blob, p)
    s.advanceTime(0.05, 40)
    b = s.bodies()
    nb = one.a["body_type"].shape[0]
    x = b["x"].reshape(copies, nb, 3)
    assert np.abs(x - x[0]).max() < 1e-3
    c = s.contacts()
    assert (c["body1"] // nb == c["body2"] // nb).all()  # no contact crosses scenes
    counts = np.bincount(c["body1"] // nb, minlength=copies)
    assert counts.min() > 0 and counts.max() - counts.min() <= 0.05 * counts.max()
    s.close()
