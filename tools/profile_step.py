"""Tiny driver for ncu: a few steps of the box-stack workload (no timing, no oracle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptivemerging_b200.ctypes_defs import default_params
from adaptivemerging_b200.scene import box_stack
from adaptivemerging_b200.system import RigidBodySystem

size = int(sys.argv[1]) if len(sys.argv) > 1 else 40
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
merging = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = default_params()
p.enable_merging = merging
s = RigidBodySystem(0).load(box_stack(size, size, size), p)
for _ in range(steps):
    s.advanceTime(0.05)
t = s.timings()
print("contacts", t.n_contacts, "pairs", t.n_pairs, "colors", t.pgs_colors, "iters", t.pgs_iterations, "step ms", t.compute_time * 1e3)
