"""Shared helpers for the parity tests."""
import numpy as np

from adaptivemerging_b200.ctypes_defs import contact_keys, default_params
from adaptivemerging_b200.scene import SceneBuilder, box_stack


def params(**kw):
    p = default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def key_index(arr):
    """dict full-key tuple -> list of indices (box x tree leaf hits share the warm-start key, the leaf
    field makes the full key unique)."""
    d = {}
    for i, k in enumerate(map(tuple, contact_keys(arr).tolist())):
        d.setdefault(k, []).append(i)
    return d


def assert_contact_sets_equal(cg, co, exact=True, tol=0.0):
    """Bit-exact comparison of two contact sets (GPU canonical order vs oracle emission order)."""
    kg, ko = key_index(cg), key_index(co)
    assert set(kg) == set(ko), (f"contact keys differ: only gpu {sorted(set(kg) - set(ko))[:5]} "
                                f"only oracle {sorted(set(ko) - set(kg))[:5]} ({len(cg)} vs {len(co)})")
    ig = np.array([kg[k][0] for k in kg])
    io = np.array([ko[k][0] for k in kg])
    assert all(len(v) == 1 for v in kg.values()) and all(len(v) == 1 for v in ko.values())
    for f in ("point_w", "normal_w", "violation"):
        a, b = cg[f][ig], co[f][io]
        if exact:
            assert np.array_equal(a, b), f"{f} differs: max |d| = {np.abs(a - b).max():.3e}"
        else:
            assert np.abs(a - b).max() <= tol, f"{f} differs: {np.abs(a - b).max():.3e}"
    return ig, io


def small_pile(nx=3, ny=4, nz=3):
    return box_stack(nx, ny, nz, pile=True)


def mixed_scene():
    """boxes, spheres, a plane, a pinned box and a spring-free composite-less mix for detection tests."""
    sb = SceneBuilder()
    sb.add_plane((0, 0, 0), (0, 1, 0))
    sb.add_box((6, 1, 6), (0, 0.5, 0), pinned=True, name="slab")
    k = 0
    for i in range(3):
        for j in range(3):
            y = 1.6 + 1.05 * ((i + j) % 3)
            if (i + j) % 2 == 0:
                sb.add_box((1, 1, 1), (-1.2 + 1.2 * i, y, -1.2 + 1.2 * j), axis_angle=(0, 1, 0, 0.3 * k), name=f"b{k}")
            else:
                sb.add_sphere(0.5, (-1.2 + 1.2 * i, y, -1.2 + 1.2 * j), name=f"s{k}")
            k += 1
    sb.add_sphere(0.6, (0.1, 3.9, 0.1), v=(0, -1, 0), name="top")
    return sb.build()


def pile_with_bullet(nx=2, ny=3, nz=2, height=45.0):
    """a small pile that merges with the plane, and one box dropped from high above that hits it later"""
    from adaptivemerging_b200.scene import _boxes_blob
    base = box_stack(nx, ny, nz, pile=True)
    xs = base.a["body_x"][1:].copy()
    Rs = base.a["body_R"][1:].reshape(-1, 3, 3).copy()
    xs = np.concatenate([xs, [[0.35, height, 0.25]]])
    Rs = np.concatenate([Rs, np.eye(3)[None]])
    return _boxes_blob(np.ones((len(xs), 3)), xs, Rs)


def lockstep(blob, p, steps, poke=None, tol=1e-6, options=None, script=None):
    """script: {step: f(system)} applied to BOTH sides before that step (UI / Factory hooks)"""
    from adaptivemerging_b200.system import RigidBodySystem
    from oracle.oracle import Oracle
    gpu = RigidBodySystem(0).load(blob, p)
    cpu = Oracle(blob, p)
    gpu.record_orders(True)
    for k, v in (options or {}).items():
        gpu.set_option(k, v)
    worst = 0.0
    gpu.hub_contacts = 0
    for step in range(steps):
        if script and step in script:
            script[step](gpu)
            script[step](cpu)
        if poke and step in poke:
            body, dv, dw = poke[step]
            gpu.add_body_velocity(body, dv, dw)
            cpu.add_body_velocity(body, dv, dw)
        gpu.advanceTime(0.05)
        full, sweep = gpu.order(0), gpu.order(1)
        post = gpu.order(2) if p.enable_post_stabilization else []  # the position-level solve of postStabilization
        gpu.hub_contacts += int((full["hub_mask"] != 0).sum()) + int((sweep["hub_mask"] != 0).sum())
        cpu.set_next_orders(full=full if len(full) else None, sweep=sweep if len(sweep) else None, post=post if len(post) else None)
        mism = cpu.step(0.05)
        assert mism == 0, f"step {step}: contact lists diverged ({mism} mismatches)"
        g, o = gpu.bodies(), cpu.bodies()
        err = max(np.abs(g["x"] - o["x"]).max(), np.abs(g["v"] - o["v"]).max(), np.abs(g["R"] - o["R"]).max())
        worst = max(worst, err)
        gpu.max_contacts = max(getattr(gpu, "max_contacts", 0), gpu.timings().n_contacts)
        assert gpu.timings().n_contacts == cpu.timings().n_contacts, f"step {step}: contact counts differ"
        assert gpu.timings().n_bodies == cpu.timings().n_bodies, f"step {step}: bodies.size() differs"  # CSV column 1
        assert err < tol, f"step {step}: state error {err:.3e}"
        assert np.array_equal(g["sleeping"], o["sleeping"]), f"step {step}: sleeping flags differ"
        assert np.array_equal(g["collection"] >= 0, o["collection"] >= 0), f"step {step}: merged sets differ"
        # RigidBodySystem.bodies list order (decides Contact.body1/body2 and the emission order of later steps)
        lo = cpu.list_order()
        inlist = lo >= 0  # (dormant bodies are in no list)
        assert np.array_equal(np.unique(gpu.list_order(raw=True)[inlist], return_inverse=True)[1],
                              np.unique(lo[inlist], return_inverse=True)[1]), f"step {step}: body list order differs"
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




GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")


def golden_scene(name):
    from adaptivemerging_b200.scene import load_blob
    import os
    return load_blob(os.path.join(GOLDEN, f"scene_{name}.npz"))


def golden_oracle(name):
    import os
    return np.load(os.path.join(GOLDEN, f"oracle_{name}.npz"))


def hub_scene():
    """a heavy unpinned plate on the plane carrying a 4x4 grid of small two-box towers: the plate is a hub of the
    contact graph (16 body pairs), the small boxes are not"""
    sb = SceneBuilder()
    sb.add_plane((0, 0, 0), (0, 1, 0))
    sb.add_box((8, 1, 8), (0, 0.5, 0), name="plate")
    k = 0
    for i in range(4):
        for j in range(4):
            x, z = -3 + 2 * i, -3 + 2 * j
            sb.add_box((1, 1, 1), (x, 1.5, z), axis_angle=(0, 1, 0, 0.1 * k), name=f"a{k}")
            sb.add_box((0.8, 0.8, 0.8), (x + 0.05, 2.4, z - 0.05), name=f"b{k}")
            k += 1
    return sb.build()
