"""Locate where the GPU and the order-replaying oracle part ways on config D (dominosPlatforms.xml, push on domino66
at step 600): prints the first step whose contact lists, contact counts, merged sets or states differ, with the
offending contacts, and dumps both sides' contacts / bodies at that step to gpurun_out/config_d_divergence.npz.
Usage (GPU box): python tools/debug_config_d.py [steps=1000]"""
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
prev = None


def keyset(c):
    return set(map(tuple, contact_keys(c).tolist()))


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
    cgi, coi = gpu.contacts(True), cpu.contacts(True)
    kgi, koi = keyset(cgi), keyset(coi)
    lg, lo = gpu.list_order(), np.unique(cpu.list_order(), return_inverse=True)[1]
    if not np.array_equal(lg, lo) and not globals().get("order_reported"):
        order_reported = True
        d = np.nonzero(lg != lo)[0]
        print(f"step {step}: list order differs at bodies {d[:20]} gpu ranks {lg[d[:20]]} oracle ranks {lo[d[:20]]}")
        print("  collections gpu", g["collection"][d[:20]], "oracle", o["collection"][d[:20]])
        eg = gpu.events(); eo = cpu.events()
        print("  events this step (gpu):", [tuple(e) for e in eg.tolist() if e[0] >= step - 1][:40])
        print("  events this step (ora):", [tuple(e) for e in eo.tolist() if e[0] >= step - 1][:40])
        np.savez("gpurun_out/config_d_listorder.npz", step=step, lg=lg, lo=lo, gcol=g["collection"], ocol=o["collection"],
                 plg=prev_l[0], plo=prev_l[1], pgcol=prev[0]["collection"], pocol=prev[1]["collection"], eg=eg, eo=eo)
    prev_l = (lg, lo)
    bad = mism != 0 or len(cg) != len(co) or not np.array_equal(g["collection"] >= 0, o["collection"] >= 0) or \
        not np.array_equal(g["sleeping"], o["sleeping"]) or kgi != koi
    if step % 50 == 0 or bad:
        print(f"step {step}: contacts gpu {len(cg)} / oracle {len(co)}, with internal {len(cgi)} / {len(coi)}, order mismatches {mism}, "
              f"state err {err:.3e}, top-level gpu {gpu.timings().n_bodies} / oracle {cpu.timings().n_bodies}, "
              f"events gpu {len(gpu.events())} / oracle {len(cpu.events())}", flush=True)
    if bad:
        kg, ko = keyset(cg), keyset(co)
        print("external only on the GPU :", sorted(kg - ko)[:20])
        print("external only in oracle  :", sorted(ko - kg)[:20])
        print("ext+int only on the GPU  :", sorted(kgi - koi)[:20])
        print("ext+int only in oracle   :", sorted(koi - kgi)[:20])
        kf, ks = keyset(full), keyset(sweep)
        print("full order keys not in oracle's external set:", sorted(kf - ko)[:20], "| missing from it:", sorted(ko - kf)[:20])
        print("sweep order len", len(sweep), "| sweep keys not in oracle ext+int:", sorted(ks - koi)[:20])
        worst = int(np.abs(g["x"] - o["x"]).max(axis=1).argmax())
        print("largest position difference: body", worst, blob.names[worst], g["x"][worst], o["x"][worst])
        print("merged-set differences at bodies:", np.nonzero((g["collection"] >= 0) != (o["collection"] >= 0))[0][:20])
        print("sleeping differences at bodies  :", np.nonzero(g["sleeping"] != o["sleeping"])[0][:20])
        eg, eo = gpu.events(), cpu.events()
        sg, so = set(map(tuple, eg.tolist())), set(map(tuple, eo.tolist()))
        print("events only gpu:", sorted(sg - so)[:20], "only oracle:", sorted(so - sg)[:20])
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez("gpurun_out/config_d_divergence.npz", step=step, cg=cgi, co=coi, full=full, sweep=sweep,
                 ev_g=eg, ev_o=eo, **{f"g_{k}": v for k, v in g.items()}, **{f"o_{k}": v for k, v in o.items()},
                 **({f"pg_{k}": v for k, v in prev[0].items()} if prev else {}),
                 **({f"po_{k}": v for k, v in prev[1].items()} if prev else {}),
                 gbpc=gpu.bpcs(), obpc=cpu.bpcs(), gibpc=gpu.internal_bpcs(), obpc_all=cpu.bpcs(True))
        break
    prev = (g, o)
else:
    print("no divergence in", steps, "steps")
