"""GPU parity with adaptive merging on: merge / unmerge decisions and trajectories against the oracle.

The oracle replays the GPU's Gauss-Seidel sequences (full solve and single sweep).  Decisions must be
IDENTICAL over the stated horizon; body states agree to 1e-6 (collection mass properties are summed in a
different floating-point order on the GPU, SURVEY.md "Hard parts" 7, and sin/cos differ by ulps).
"""
import numpy as np
import pytest

from tests.util import mixed_scene, params, small_pile

pytestmark = pytest.mark.gpu


def lockstep(blob, p, steps, poke=None, tol=1e-6):
    from adaptivemerging_b200.system import RigidBodySystem
    from oracle.oracle import Oracle
    gpu = RigidBodySystem(0).load(blob, p)
    cpu = Oracle(blob, p)
    gpu.record_orders(True)
    worst = 0.0
    for step in range(steps):
        if poke and step in poke:
            body, dv, dw = poke[step]
            gpu.add_body_velocity(body, dv, dw)
            cpu.add_body_velocity(body, dv, dw)
        gpu.advanceTime(0.05)
        full, sweep = gpu.order(0), gpu.order(1)
        cpu.set_next_orders(full=full if len(full) else None, sweep=sweep if len(sweep) else None)
        mism = cpu.step(0.05)
        assert mism == 0, f"step {step}: contact lists diverged ({mism} mismatches)"
        g, o = gpu.bodies(), cpu.bodies()
        err = max(np.abs(g["x"] - o["x"]).max(), np.abs(g["v"] - o["v"]).max(), np.abs(g["R"] - o["R"]).max())
        worst = max(worst, err)
        assert err < tol, f"step {step}: state error {err:.3e}"
        assert np.array_equal(g["sleeping"], o["sleeping"]), f"step {step}: sleeping flags differ"
        assert np.array_equal(g["collection"] >= 0, o["collection"] >= 0), f"step {step}: merged sets differ"
    ev_g = sorted(map(tuple, gpu.events().tolist()))
    ev_o = sorted(map(tuple, cpu.events().tolist()))
    return gpu, cpu, ev_g, ev_o, worst


def same_partition(a, b):
    """collection ids are arbitrary: compare the induced partitions"""
    ma, mb = {}, {}
    for x, y in zip(a.tolist(), b.tolist()):
        if (x < 0) != (y < 0):
            return False
        if x >= 0:
            if ma.setdefault(x, y) != y or mb.setdefault(y, x) != x:
                return False
    return True


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
