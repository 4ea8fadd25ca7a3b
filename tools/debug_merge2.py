import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import small_pile, params, mixed_scene
from adaptivemerging_b200.ctypes_defs import contact_keys
from adaptivemerging_b200.system import RigidBodySystem
from oracle.oracle import Oracle
from tests.util import pile_with_bullet
blob = pile_with_bullet() if len(sys.argv) > 2 else small_pile()
p = params()
gpu = RigidBodySystem(0).load(blob, p); cpu = Oracle(blob, p)
gpu.record_orders(True)
import os
if os.environ.get('HUB'): gpu.set_option('hub_min_degree', int(os.environ['HUB']))
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 14):
    gpu.advanceTime(0.05)
    full, sweep = gpu.order(0), gpu.order(1)
    cpu.set_next_orders(full=full if len(full) else None, sweep=sweep if len(sweep) else None)
    mism = cpu.step(0.05)
    g, o = gpu.bodies(), cpu.bodies()
    errs = {k: float(np.abs(g[k] - o[k]).max()) for k in ("x", "R", "v", "omega")}
    tg, to = gpu.timings(), cpu.timings()
    print(step, "mism", mism, {k: "%.2e" % v for k, v in errs.items()}, "contacts", tg.n_contacts, to.n_contacts, "ncoll", tg.n_collections, to.n_collections,
          "events", len(gpu.events()), len(cpu.events()), "sleep", g["sleeping"].sum(), o["sleeping"].sum(), "iters", tg.pgs_iterations, to.pgs_iterations, "nsweep", len(sweep), "nfull", len(full))
    if step >= 9:
        ib = gpu.internal_bpcs(); ob = cpu.bpcs(True); ob = ob[ob["in_collection"] != 0]
        gi = sorted((min(a, b), max(a, b), n, m) for a, b, n, m in zip(ib["body1"].tolist(), ib["body2"].tolist(), ib["n_contacts"].tolist(), ib["n_metric"].tolist()))
        oi = sorted((min(a, b), max(a, b), n, m) for a, b, n, m in zip(ob["body1"].tolist(), ob["body2"].tolist(), ob["n_contacts"].tolist(), ob["n_metric"].tolist()))
        print("  internal gpu", gi); print("  internal cpu", oi)
    if max(errs.values()) > 1e-9 and not mism:
        cg, co = gpu.contacts(), cpu.contacts()
        from tests.util import key_index
        ko = key_index(co)
        io = np.array([ko[k][0] for k in map(tuple, contact_keys(cg).tolist())])
        dl = np.abs(cg["lambda"] - co["lambda"][io]).max(1)
        print(" lambda max diff", dl.max(), "at", int(dl.argmax()), cg[int(dl.argmax())][["body1","body2","info"]], "warm diff", np.abs(cg["lambda_warm"] - co["lambda_warm"][io]).max(),
              "iters", tg.pgs_iterations, to.pgs_iterations, "hub contacts", int((full["hub_mask"] != 0).sum()), "colors", tg.pgs_colors)
        break
    if mism:
        kg = set(map(tuple, contact_keys(gpu.contacts(True)).tolist()))
        ko = set(map(tuple, contact_keys(cpu.contacts(True)).tolist()))
        print(" only gpu", sorted(kg - ko)[:12]); print(" only cpu", sorted(ko - kg)[:12])
        ks = set(map(tuple, contact_keys(sweep).tolist())); kf = set(map(tuple, contact_keys(full).tolist()))
        print(" sweep not in cpu", sorted(ks - ko)[:8], " full not in cpu", sorted(kf - ko)[:8])
        print(" gpu coll", g["collection"].tolist()); print(" cpu coll", o["collection"].tolist())
        print(" gpu events", gpu.events()[-8:].tolist()); print(" cpu events", cpu.events()[-8:].tolist())
        break
