import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import small_pile, params, mixed_scene
from adaptivemerging_b200.system import RigidBodySystem
from oracle.oracle import Oracle
blob = small_pile()
p = params()
gpu = RigidBodySystem(0).load(blob, p); cpu = Oracle(blob, p)
gpu.record_orders(True)
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 14):
    gpu.advanceTime(0.05)
    full, sweep = gpu.order(0), gpu.order(1)
    cpu.set_next_orders(full=full if len(full) else None, sweep=sweep if len(sweep) else None)
    mism = cpu.step(0.05)
    g, o = gpu.bodies(), cpu.bodies()
    errs = {k: float(np.abs(g[k] - o[k]).max()) for k in ("x", "R", "v", "omega")}
    tg, to = gpu.timings(), cpu.timings()
    print(step, "mism", mism, errs, "contacts", tg.n_contacts, to.n_contacts, "ncoll", tg.n_collections, to.n_collections,
          "events", len(gpu.events()), len(cpu.events()), "sleep", g["sleeping"].sum(), o["sleeping"].sum(), "iters", tg.pgs_iterations, to.pgs_iterations)
    if max(errs.values()) > 1e-6:
        i = int(np.abs(g["v"] - o["v"]).max(1).argmax())
        print(" worst body", i, "gpu v", g["v"][i], "cpu v", o["v"][i], "coll", g["collection"][i], o["collection"][i], "x", g["x"][i], o["x"][i])
        print(" gpu events", gpu.events()[-6:].tolist()); print(" cpu events", cpu.events()[-6:].tolist())
        cg, co = gpu.contacts(), cpu.contacts()
        for nm, cc in (("gpu", cg), ("cpu", co)):
            sel = (cc["body1"] == i) | (cc["body2"] == i)
            print(nm, "contacts of worst body:")
            for r in cc[sel]:
                print("   ", r["body1"], r["body2"], r["info"], "lam", r["lambda"], "warm", r["lambda_warm"], "viol", r["violation"], "col", r["color"], "new", r["new_this_step"], "state", r["state"])
        print("gpu dv", gpu.deltav()[i], "cpu dv", cpu.deltav()[i])
        print("collections of neighbours gpu", g["collection"][[r for r in set(cg[(cg["body1"] == i) | (cg["body2"] == i)]["body1"].tolist())]])
        break
import ctypes as C
cg = gpu.contacts()
for k in np.nonzero((cg["body1"] == 36) | (cg["body2"] == 36))[0][:2]:
    out = np.zeros(50)
    gpu._L.am3d_debug_solve_row.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    rc = gpu._L.am3d_debug_solve_row(gpu._h, int(k), out.ctypes.data_as(C.c_void_p))
    print("row", k, rc, "dirs", out[:9], "r", out[9:15], "b", out[15:18], "D", out[18:21], "lam", out[21:24], "\n mass", out[24:44], "meta", out[44:])
for slot in range(0):
    print("collection", slot, gpu.collection(slot))
