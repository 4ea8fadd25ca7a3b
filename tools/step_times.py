"""Per-step phase timings of a bench workload (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from adaptivemerging_b200.system import RigidBodySystem
name = sys.argv[1]; size = int(sys.argv[2]); steps = int(sys.argv[3]); merging = int(sys.argv[4]) if len(sys.argv) > 4 else 1
every = int(sys.argv[5]) if len(sys.argv) > 5 else 1
blob, p, desc = bench.build_workload(name, size, merging, float(os.environ['Y0']) if os.environ.get('Y0') else None)
s = RigidBodySystem(0).load(blob, p)
import time
for i in range(steps):
    t0 = time.perf_counter(); s.advanceTime(0.05); w = (time.perf_counter() - t0) * 1e3
    t = s.timings()
    if (i + 1) % every == 0 or t.compute_time * 1e3 > 50:
        print(f"step {i+1:4d} wall {w:8.2f} total {t.compute_time*1e3:8.2f} det {t.detection*1e3:7.2f} warm {t.warmstart*1e3:6.2f} coll {t.update_collections*1e3:7.2f} "
              f"unm {t.unmerging*1e3:7.2f} lcp {t.lcp_solve*1e3:7.2f} sweeps {t.pgs_kernel_time*1e3:7.2f} merge {t.merging*1e3:7.2f} | contacts {t.n_contacts} pairs {t.n_pairs} "
              f"colors {t.pgs_colors} iters {t.pgs_iterations} ncoll {t.n_collections}", flush=True)
