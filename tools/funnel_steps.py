"""Per-step phase times of the funnel pile (config F at reduced size): python tools/funnel_steps.py [n] [steps] [opt=value ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptivemerging_b200.ctypes_defs import apply_overrides, default_params
from adaptivemerging_b200.scene import funnel_pile
from adaptivemerging_b200.system import RigidBodySystem
from tests.util import golden_scene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 45
blob = funnel_pile(golden_scene("funnel_template"), n, 10, n, y0=0.6)
p = apply_overrides(default_params(), blob.overrides)
s = RigidBodySystem(0).load(blob, p)
s.set_option("record_events", 0)
for kv in sys.argv[3:]:
    s.set_option(kv.split("=")[0], float(kv.split("=")[1]))
for k in range(steps):
    t0 = time.perf_counter()
    s.advanceTime(0.05)
    w = (time.perf_counter() - t0) * 1e3
    t = s.timings()
    print(f"step {k:3d} wall {w:8.1f} ms  contacts {t.n_contacts:8d} pairs {t.n_pairs:6d} detect {t.detection*1e3:7.1f} narrow {t.narrowphase_kernel_time*1e3:6.1f} "
          f"warm {t.warmstart*1e3:6.1f} lcp {t.lcp_solve*1e3:7.1f} pgs {t.pgs_kernel_time*1e3:7.1f} compute {t.compute_time*1e3:7.1f} giants {t.pgs_giant_groups}", flush=True)
