"""Locate where the GPU and the order-replaying oracle part ways on config D (dominosPlatforms.xml, push on domino66
at step 600): prints the first step whose contact lists, contact counts, merged sets or states differ, with the
offending contacts.  Usage (GPU box): python tools/debug_config_d.py [steps=1000]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from adaptivemerging_b200.ctypes_defs import apply_overrides, contact_keys, default_params  # noqa: E402
from adaptivemerging_b200.system import RigidBodySystem  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from tests.util import golden_scene  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
blob = golden_scene("dominosPlatforms")
p = apply_overrides(default_params(), blob.overrides)
gpu, cpu = RigidBodySystem(0).load(blob, p), Oracle(blob, p)
gpu.record_orders(True)
push = blob.names.index("domino66")
for step in range(steps):
    if step == 600:
        gpu.add_body_velocity(push, None, np.array([0.0, 0.0, -2.0]))
        cpu.add_body_velocity(push, None, np.array([0.0, 0.0, -2.0]))
    gpu.advanceTime(0.05)
    full, sweep = gpu.order(0), gpu.order(1)
    cpu.set_next_orders(full=full if len(full) else None, sweep=sweep if len(sweep) else None)
    mism = cpu.step(0.05)
    g, o = gpu.bodies(), cpu.bodies()
    err = max(np.abs(g["x"] - o["x"]).max(), np.abs(g["v"] - o["v"]).max(), np.abs(g["R"] - o["R"]).max())
    cg, co = gpu.contacts(), cpu.contacts()
    bad = mism != 0 or len(cg) != len(co) or not np.array_equal(g["collection"] >= 0, o["collection"] >= 0) or \
        not np.array_equal(g["sleeping"], o["sleeping"])
    if step % 50 == 0 or bad:
        print(f"step {step}: contacts gpu {len(cg)} / oracle {len(co)}, order mismatches {mism}, state err {err:.3e}, "
              f"top-level gpu {gpu.timings().n_bodies} / oracle {cpu.timings().n_bodies}", flush=True)
    if bad:
        kg = set(map(tuple, contact_keys(cg).tolist()))
        ko = set(map(tuple, contact_keys(co).tolist()))
        print("only on the GPU :", sorted(kg - ko)[:20])
        print("only in oracle  :", sorted(ko - kg)[:20])
        worst = int(np.abs(g["x"] - o["x"]).max(axis=1).argmax())
        print("largest position difference: body", worst, blob.names[worst], g["x"][worst], o["x"][worst])
        print("merged-set differences at bodies:", np.nonzero((g["collection"] >= 0) != (o["collection"] >= 0))[0][:20])
        print("sleeping differences at bodies  :", np.nonzero(g["sleeping"] != o["sleeping"])[0][:20])
        break
else:
    print("no divergence in", steps, "steps")
