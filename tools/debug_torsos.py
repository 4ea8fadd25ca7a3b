"""torsos lockstep, error per step (python tools/debug_torsos.py [opt=value ...])"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from adaptivemerging_b200.ctypes_defs import apply_overrides, default_params
from adaptivemerging_b200.system import RigidBodySystem
from oracle.oracle import Oracle
from tests.util import golden_scene
blob = golden_scene("torsos")
p = apply_overrides(default_params(), blob.overrides)
gpu = RigidBodySystem(0).load(blob, p)
cpu = Oracle(blob, p)
gpu.record_orders(True)
for kv in sys.argv[1:]:
    gpu.set_option(kv.split("=")[0], float(kv.split("=")[1]))
for step in range(125):
    gpu.advanceTime(0.05)
    full, sweep = gpu.order(0), gpu.order(1)
    cpu.set_next_orders(full=full if len(full) else None, sweep=sweep if len(sweep) else None)
    mism = cpu.step(0.05)
    g, o = gpu.bodies(), cpu.bodies()
    ex = np.abs(g["x"] - o["x"]).max(); ev = np.abs(g["v"] - o["v"]).max()
    if 108 <= step <= 112 and "-v" in os.environ.get("DBG", "-v"):
        from adaptivemerging_b200.ctypes_defs import contact_keys
        dvb = np.abs(g["v"] - o["v"]).max(1); dwb = np.abs(g["omega"] - o["omega"]).max(1)
        bad = np.nonzero((dvb > 1e-9) | (dwb > 1e-9))[0]
        print("   bodies off:", bad.tolist(), "dv", dvb[bad], "dw", dwb[bad])
        print("   collection gpu", g["collection"].tolist(), "cpu", o["collection"].tolist())
        print("   sleeping gpu", g["sleeping"].tolist(), "cpu", o["sleeping"].tolist())
        print("   events gpu", gpu.events().tolist()[-6:], "cpu", cpu.events().tolist()[-6:])
        for inc in (False, True):
            cg, co = gpu.contacts(inc), cpu.contacts(inc)
            kg = {tuple(k): i for i, k in enumerate(contact_keys(cg).tolist())}
            ko = {tuple(k): i for i, k in enumerate(contact_keys(co).tolist())}
            common = [k for k in kg if k in ko]
            dl = [(k[:2], k[4:6], np.abs(cg["lambda"][kg[k]] - co["lambda"][ko[k]]).max(), cg["lambda"][kg[k]].tolist(), co["lambda"][ko[k]].tolist(), int(cg["in_collection"][kg[k]]), int(cg["state"][kg[k]]), int(co["state"][ko[k]])) for k in common]
            dl = [d for d in dl if d[2] > 1e-12]
            print(f"   contacts(internal={inc}) gpu {len(cg)} cpu {len(co)} common {len(common)} lambda-off {len(dl)}")
            for d in dl[:12]:
                print("      ", d)
        print("   order sizes full", len(full), "sweep", len(sweep), "ncoll", gpu.timings().n_collections, "iters", gpu.timings().pgs_iterations, cpu.timings().pgs_iterations)
    print(f"step {step:3d} mism {mism} ncon {gpu.timings().n_contacts} {cpu.timings().n_contacts} err x {ex:.3e} v {ev:.3e} giants {gpu.timings().pgs_giant_groups} events {len(gpu.events())} {len(cpu.events())}", flush=True)
