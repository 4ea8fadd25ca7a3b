"""Contact-graph statistics of config F (funnel + torso pile) at one step: contacts per body pair, per-body contact load,
phases of the full solve and the longest chain per phase.  Usage (GPU box): python tools/funnel_stats.py [n=20] [steps=38]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bench import build_workload  # noqa: E402
from adaptivemerging_b200.system import RigidBodySystem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 38
blob, p, desc = build_workload("funnel", n, 1, 0.6)
s = RigidBodySystem(0).load(blob, p)
s.set_option("record_events", 0)
s.advanceTime(0.05, steps - 1)
s.record_orders(True)
s.advanceTime(0.05)
c = s.contacts()
t = s.timings()
print(desc)
print(f"step {steps}: {len(c)} contacts, {t.n_pairs} candidate pairs, {t.pgs_colors} phases, {t.pgs_iterations} iterations, "
      f"lcp {t.lcp_solve * 1e3:.1f} ms, sweeps {t.pgs_kernel_time * 1e3:.1f} ms, detection {t.detection * 1e3:.1f} ms")
pair = c["body1"].astype(np.int64) * (1 << 24) + c["body2"]
u, cnt = np.unique(pair, return_counts=True)
print(f"body pairs {len(u)}; contacts per pair: mean {cnt.mean():.1f}, median {np.median(cnt):.0f}, p99 {np.percentile(cnt, 99):.0f}, max {cnt.max()}; "
      f"pairs >= 65: {(cnt >= 65).sum()} holding {cnt[cnt >= 65].sum()} contacts")
nb = blob.n_bodies
load = np.bincount(c["body1"], minlength=nb) + np.bincount(c["body2"], minlength=nb)
pinned = (blob.a["body_flags"] & 1) != 0
pinned |= blob.a["body_type"] == 1
free = load.copy()
free[pinned] = 0
top = np.argsort(-free)[:8]
print("largest per-body contact loads (unpinned):", [(int(b), blob.names[b] if blob.names else "", int(free[b])) for b in top])
print("largest loads incl. pinned:", [(int(b), int(load[b])) for b in np.argsort(-load)[:5]])
deg = np.bincount(u >> 24, minlength=nb) + np.bincount(u & 0xffffff, minlength=nb)
print("pairs per body: max unpinned", int(deg[~pinned].max()), "max pinned", int(deg[pinned].max()) if pinned.any() else 0)
o = s.order(0)
col = o["color"]
ph, first = np.unique(col, return_index=True)
# longest chain per phase = max contacts of one (body1, body2, phase) run
key = o["body1"].astype(np.int64) * (1 << 24) + o["body2"]
chain = []
for k in ph:
    m = col == k
    _, cc = np.unique(key[m], return_counts=True)
    chain.append(cc.max())
chain = np.array(chain)
print(f"phases {len(ph)}; sum over phases of the longest chain: {chain.sum()} contacts; phases with a chain >= 65: {(chain >= 65).sum()}; "
      f"hub contacts {(o['hub_mask'] != 0).sum()}")
print("per-phase longest chain:", chain.tolist()[:80])
