"""Per-step phase times of N batched copies of tower25platform.xml: python tools/batch_steps.py [copies] [steps] [opt=value ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload
from adaptivemerging_b200.system import RigidBodySystem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 170
blob, p, desc = build_workload("batch", n, 1, None)
s = RigidBodySystem(0).load(blob, p)
s.set_option("record_events", 0)
for kv in sys.argv[3:]:
    s.set_option(kv.split("=")[0], float(kv.split("=")[1]))
prof = int(os.environ.get("AM3D_PROFILE_STEP", "-1"))   # ncu --profile-from-start off: capture exactly this step
for k in range(steps):
    if k == prof:
        import torch
        torch.cuda.cudart().cudaProfilerStart()
    t0 = time.perf_counter()
    s.advanceTime(0.05)
    w = (time.perf_counter() - t0) * 1e3
    if k == prof:
        torch.cuda.cudart().cudaProfilerStop()
    t = s.timings()
    if k >= 120:
        print(f"step {k:3d} wall {w:6.2f} contacts {t.n_contacts:8d} coll {t.n_collections:5d} detect {t.detection*1e3:5.2f} warm {t.warmstart*1e3:5.2f} "
              f"upd {t.update_collections*1e3:5.2f} (sweep {t.single_it_pgs*1e3:4.2f}) unmerge {t.unmerging*1e3:5.2f} (build {t.unmerging_build*1e3:4.2f}) lcp {t.lcp_solve*1e3:5.2f} "
              f"(pgs {t.pgs_kernel_time*1e3:5.2f}) merge {t.merging*1e3:5.2f} total {t.compute_time*1e3:6.2f} phases {t.pgs_colors}", flush=True)
