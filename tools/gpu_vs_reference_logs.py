"""Longest prefix of the reference's recorded (#bodies, #contacts) series (tests/golden/ref_logs_tower25platform.npz)
that the GPU step reproduces exactly (free running, colour-ordered Gauss-Seidel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from adaptivemerging_b200.ctypes_defs import apply_overrides, default_params
from adaptivemerging_b200.system import RigidBodySystem
from tests.util import golden_scene
from tests.test_reference_logs import CASES, GOLDEN
for name, (ov, rows) in sorted(CASES.items()):
    ref = np.load(os.path.join(GOLDEN, "ref_logs_tower25platform.npz"))[name]
    blob = golden_scene("tower25platform")
    p = apply_overrides(default_params(), blob.overrides)
    for k, v in ov.items(): setattr(p, k, v)
    s = RigidBodySystem(0).load(blob, p)
    s.advanceTime(0.05)
    mine = []
    for _ in range(130):
        s.advanceTime(0.05); t = s.timings(); mine.append((t.n_bodies, t.n_contacts))
    mine = np.array(mine)
    bad = np.nonzero((mine != ref[:130]).any(1))[0]
    m = bad[0] if len(bad) else 130
    print(name, "gpu matches", m, "rows (oracle:", rows, ")", "ref", ref[min(m, 129)], "gpu", mine[min(m, 129)], flush=True)
    s.close()
