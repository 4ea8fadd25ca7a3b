"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle.

Tolerances: contact sets (keys, points, normals, depths) are compared BIT-EXACTLY from identical input
state (teacher forcing); PGS impulses and body deltaV are compared bit-exactly when the oracle replays the
GPU's colour order; free-running trajectories are compared to 1e-9 (sin/cos of the rotation update are
not bit-identical between glibc/Java and CUDA, SURVEY.md "Hard parts" 1).
"""
import numpy as np
import pytest

from tests.util import assert_contact_sets_equal, key_index, mixed_scene, params, small_pile

pytestmark = pytest.mark.gpu


def _pair(blob, **kw):
    from adaptivemerging_b200.system import RigidBodySystem
    from oracle.oracle import Oracle
    p = params(**kw)
    return RigidBodySystem(0).load(blob, p), Oracle(blob, p)


@pytest.mark.parametrize("scene", ["pile", "mixed"])
def test_contact_sets_bit_exact_teacher_forced(scene):
    blob = small_pile() if scene == "pile" else mixed_scene()
    gpu, cpu = _pair(blob, enable_merging=0)
    checked = 0
    for step in range(60):
        cpu.step(0.05)
        if step % 5 != 4:
            continue
        b = cpu.bodies()
        gpu.upload_bodies(b["x"], b["R"], b["v"], b["omega"])
        ng = gpu.detect()
        no = cpu.detect()
        assert ng == no
        if no:
            assert_contact_sets_equal(gpu.contacts(), cpu.contacts())
            checked += no
    assert checked > 50


@pytest.mark.parametrize("scene", ["torsos", "funnel"])
def test_contact_sets_bit_exact_teacher_forced_trees_and_composites(scene):
    """the same for sphere trees against the plane / each other (torso_flux.sph meshes) and for composite parts against
    trees and boxes (funnel.xml): identity incl. leaf ids and part indices, world points, normals, depths, bit for bit"""
    from adaptivemerging_b200.ctypes_defs import apply_overrides
    from adaptivemerging_b200.scene import funnel_pile
    from tests.util import golden_scene
    if scene == "torsos":
        blob, steps, every = golden_scene("torsos"), 90, 6
    else:
        blob, steps, every = funnel_pile(golden_scene("funnel_template"), nx=3, ny=2, nz=3, pitch=6.0, y0=52.0), 150, 10
    from adaptivemerging_b200.system import RigidBodySystem
    from oracle.oracle import Oracle
    p = apply_overrides(params(enable_merging=0), blob.overrides)
    p.enable_merging = 0
    gpu, cpu = RigidBodySystem(0).load(blob, p), Oracle(blob, p)
    checked = 0
    for step in range(steps):
        cpu.step(0.05)
        if step % every != every - 1:
            continue
        b = cpu.bodies()
        gpu.upload_bodies(b["x"], b["R"], b["v"], b["omega"])
        ng, no = gpu.detect(), cpu.detect()
        assert ng == no
        if no:
            assert_contact_sets_equal(gpu.contacts(), cpu.contacts())
            checked += no
    assert checked > 200


def test_hub_solve_residuals_against_the_reference_order():
    """Hub bodies are the one place where the GPU's sequence is not a permutation of the reference's (Jacobi across the
    groups of a colour for the hub's deltaV).  Evidence that does not rest on the oracle's replay of that scheme: after
    the same 30 iterations the LCP residuals of the hub solve are compared with those of the reference's own list-order
    Gauss-Seidel on the same contacts."""
    from oracle.oracle import Oracle
    from tests.util import hub_scene
    blob = hub_scene()
    gpu, cpu = _pair(blob, enable_merging=0)
    ref = Oracle(blob, params(enable_merging=0))
    gpu.set_option("hub_min_degree", 8)
    gpu.record_orders(True)
    for _ in range(40):
        cpu.step(0.05)
        ref.step(0.05)
    b = cpu.bodies()
    gpu.upload_bodies(b["x"], b["R"], b["v"], b["omega"])
    assert gpu.detect() == cpu.detect() == ref.detect() > 0
    gpu.solve(0.05)
    order = gpu.order(0)
    assert (order["hub_mask"] != 0).sum() > 30
    cpu.apply_external_forces()
    assert cpu.solve(0.05, order) == 0
    cg, co = gpu.contacts(), cpu.contacts()
    ko = key_index(co)
    io = np.array([ko[k][0] for k in map(tuple, np.stack([cg[f] for f in ("body1", "body2", "csb1", "csb2", "bv1", "bv2", "info", "leaf")], 1).tolist())])
    assert np.array_equal(cg["lambda"], co["lambda"][io])           # the oracle holds the GPU's hub solution
    ref.apply_external_forces()
    ref.solve(0.05)
    r_gpu, r_ref = cpu.residuals(), ref.residuals()
    print(f"\nhub solve, {int(r_gpu[3])} contacts: GPU normal {r_gpu[0]:.3e} cone {r_gpu[1]:.3e} tangential {r_gpu[2]:.3e} | reference order "
          f"normal {r_ref[0]:.3e} cone {r_ref[1]:.3e} tangential {r_ref[2]:.3e}")
    assert r_gpu[1] <= 1e-12 and r_ref[1] <= 1e-12
    assert r_gpu[0] <= max(5 * r_ref[0], 2e-3) and r_gpu[2] <= max(5 * r_ref[2], 2e-3)
    # and the two solutions are the same physical answer
    assert np.abs(cpu.deltav() - ref.deltav()).max() < 5e-3


@pytest.mark.parametrize("hub_min", [64, 2])
def test_pgs_bit_exact_in_colour_order(hub_min):
    """hub_min = 2 forces the hub path (bodies touched by >= 2 body pairs are solved Jacobi-style across a colour,
    DESIGN.md section 4); the oracle replays that sequence too"""
    blob = small_pile(4, 5, 4)
    gpu, cpu = _pair(blob, enable_merging=0)
    gpu.set_option("hub_min_degree", hub_min)
    gpu.record_orders(True)
    for _ in range(25):
        cpu.step(0.05)
    b = cpu.bodies()
    gpu.upload_bodies(b["x"], b["R"], b["v"], b["omega"])
    assert gpu.detect() == cpu.detect() > 0
    gpu.solve(0.05)
    cg = gpu.contacts()
    order = gpu.order(0)
    assert (order["hub_mask"] != 0).any() == (hub_min == 2)
    cpu.apply_external_forces()
    mism = cpu.solve(0.05, order)
    assert mism == 0
    co = cpu.contacts()
    ko = key_index(co)
    io = np.array([ko[k][0] for k in map(tuple, np.stack([cg[f] for f in
                  ("body1", "body2", "csb1", "csb2", "bv1", "bv2", "info", "leaf")], 1).tolist())])
    assert gpu.timings().pgs_iterations == cpu.timings().pgs_iterations
    assert np.array_equal(cg["lambda"], co["lambda"][io]), np.abs(cg["lambda"] - co["lambda"][io]).max()
    assert np.array_equal(gpu.deltav(), cpu.deltav())
    assert np.array_equal(cg["state"], co["state"][io])


@pytest.mark.parametrize("scene", ["pile", "mixed"])
def test_free_running_without_merging(scene):
    blob = small_pile() if scene == "pile" else mixed_scene()
    gpu, cpu = _pair(blob, enable_merging=0)
    for step in range(40):
        gpu.advanceTime(0.05)
        # replay the GPU's Gauss-Seidel order on the oracle
        cg = gpu.contacts()
        if len(cg):
            cpu.set_next_orders(full=cg[gpu.solve_order()])
        mism = cpu.step(0.05)
        assert mism == 0, f"step {step}: contact sets diverged"
        g, o = gpu.bodies(), cpu.bodies()
        assert np.abs(g["x"] - o["x"]).max() < 1e-9, step
        assert np.abs(g["v"] - o["v"]).max() < 1e-9, step
        assert np.array_equal(g["sleeping"], o["sleeping"]), step


def test_lcp_residuals_against_unpermuted_reference():
    """North-star parity mode 3: complementarity residuals of the GPU's colour-ordered solve next to those of the
    reference's own (list-order) Gauss-Seidel on the same contacts.  w = b + J dv + c*lambda; normal rows:
    |min(lambda, w)|, friction rows: cone violation and |w_t| strictly inside the cone."""
    from oracle.oracle import Oracle
    blob = small_pile(4, 5, 4)
    gpu, cpu = _pair(blob, enable_merging=0)
    ref = Oracle(blob, params(enable_merging=0))
    gpu.record_orders(True)
    for _ in range(25):
        cpu.step(0.05)
        ref.step(0.05)
    b = cpu.bodies()
    gpu.upload_bodies(b["x"], b["R"], b["v"], b["omega"])
    assert gpu.detect() == cpu.detect() == ref.detect() > 0
    gpu.solve(0.05)
    cpu.apply_external_forces()
    assert cpu.solve(0.05, gpu.order(0)) == 0       # the oracle now holds exactly the GPU's lambda and deltaV
    ref.apply_external_forces()
    ref.solve(0.05)                                 # the reference's own order
    r_gpu, r_ref = cpu.residuals(), ref.residuals()
    print(f"\nLCP residuals after 30 sweeps, {int(r_gpu[3])} contacts: colour order (GPU) normal {r_gpu[0]:.3e} cone "
          f"{r_gpu[1]:.3e} tangential {r_gpu[2]:.3e} | reference order normal {r_ref[0]:.3e} cone {r_ref[1]:.3e} "
          f"tangential {r_ref[2]:.3e}")
    assert r_gpu[3] == r_ref[3] > 0
    assert r_gpu[1] <= 1e-12 and r_ref[1] <= 1e-12                  # projections hold exactly on both sides
    assert r_gpu[0] <= max(3 * r_ref[0], 1e-3) and r_gpu[2] <= max(3 * r_ref[2], 1e-3)
