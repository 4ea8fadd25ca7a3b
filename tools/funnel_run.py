"""Config F (funnel + n x 10 x n torso_flux sphere-tree bodies) for a number of steps: per-step timings into a JSON-lines
file.  Usage (GPU box): python tools/funnel_run.py n steps y0 out.jsonl [max_contacts]
(max_contacts: stop once a step holds more contacts than this - the full-size pile outgrows one GPU's memory while it lands)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bench import build_workload  # noqa: E402
from adaptivemerging_b200.system import RigidBodySystem  # noqa: E402

n, steps, y0, out = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
max_contacts = float(sys.argv[5]) if len(sys.argv) > 5 else float("inf")
t0 = time.time()
blob, p, desc = build_workload("funnel", n, 1, y0)
nb = int((blob.a["body_type"] != 1).sum())
s = RigidBodySystem(0).load(blob, p)
s.set_option("record_events", 0)
print(desc, f"| {nb} bodies | scene built and uploaded in {time.time() - t0:.1f} s", flush=True)
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
with open(out, "w") as f:
    f.write(json.dumps({"description": desc, "bodies": nb}) + "\n")
    wall0 = time.time()
    for k in range(steps):
        s.advanceTime(0.05)
        t = s.timings()
        st = s.stats()
        row = {"step": k + 1, "contacts": t.n_contacts, "pairs": t.n_pairs, "collections": t.n_collections, "top_level": t.n_bodies,
               "ms": t.compute_time * 1e3, "detection_ms": t.detection * 1e3, "narrowphase_ms": t.narrowphase_kernel_time * 1e3,
               "warmstart_ms": t.warmstart * 1e3, "lcp_ms": t.lcp_solve * 1e3, "pgs_sweeps_ms": t.pgs_kernel_time * 1e3,
               "update_collections_ms": t.update_collections * 1e3, "phases": t.pgs_colors, "iterations": t.pgs_iterations,
               "row_updates_total": st["row_updates"], "solve_seconds_total": st["solve_seconds"], "wall_s": time.time() - wall0}
        f.write(json.dumps(row) + "\n")
        f.flush()
        if (k + 1) % 10 == 0:
            print(row, flush=True)
        if t.n_contacts > max_contacts:
            print(f"stopping after step {k + 1}: {t.n_contacts} contacts > {max_contacts:.3g}", flush=True)
            print(row, flush=True)
            break
b = s.bodies()
print("finite:", bool(np.isfinite(b["x"]).all()), "y range", float(b["x"][:, 1].min()), float(b["x"][:, 1].max()), "total wall", time.time() - wall0)
