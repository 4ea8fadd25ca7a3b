"""Turn a `HeadlessStep --dump` file (real Java reference, baseline/HeadlessStep.java) into tests/golden/java_<scene>.npz.

    python tools/java_dump_to_npz.py dump.txt tests/golden/scene_<scene>.npz tests/golden/java_<scene>.npz [every=10]

Stored per dumped step (every `every`-th step and the last): body states in the scene blob's body order (matched by name),
top-level body count, contact count, and per contact (body1, body2, info, lambda[3], contactB1[3], normalB1[3], violation).
tests/test_java_golden.py compares the oracle with it."""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from adaptivemerging_b200.scene import load_blob  # noqa: E402


def main():
    dump, blob_path, out = sys.argv[1:4]
    every = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    blob = load_blob(blob_path)
    index = {n: i for i, n in enumerate(blob.names)}
    nb = blob.n_bodies
    steps, tops, ncs, X, R, V, W, SL, CO, contacts = [], [], [], [], [], [], [], [], [], []
    cur = None
    with open(dump) as f:
        for ln in f:
            t = ln.split()
            if t[0] == "S":
                if cur is not None:
                    flush(cur, every, steps, tops, ncs, X, R, V, W, SL, CO, contacts)
                cur = dict(step=int(t[1]), top=int(t[2]), nc=int(t[4]), x=np.full((nb, 3), np.nan), R=np.full((nb, 9), np.nan),
                           v=np.zeros((nb, 3)), w=np.zeros((nb, 3)), sl=np.zeros(nb, np.int32), co=np.full(nb, -1, np.int32), c=[])
            elif t[0] == "B":
                i = index[t[1]]
                vals = list(map(float, t[2:20]))
                cur["x"][i] = vals[0:3]; cur["R"][i] = vals[3:12]; cur["v"][i] = vals[12:15]; cur["w"][i] = vals[15:18]
                cur["sl"][i] = int(t[20]); cur["co"][i] = int(t[21])
            elif t[0] == "C":
                cur["c"].append([index[t[1]], index[t[2]], int(t[7])] + list(map(float, t[8:18])))
    if cur is not None:
        flush(cur, 1, steps, tops, ncs, X, R, V, W, SL, CO, contacts)
    off = np.cumsum([0] + [len(c) for c in contacts])
    allc = np.array([r for c in contacts for r in c], dtype=np.float64).reshape(-1, 13)
    np.savez_compressed(out, steps=np.array(steps), top_level=np.array(tops), n_contacts=np.array(ncs), x=np.array(X), R=np.array(R),
                        v=np.array(V), omega=np.array(W), sleeping=np.array(SL), collection=np.array(CO), contact_offsets=off, contacts=allc)
    print(f"{len(steps)} steps -> {out}")


def flush(cur, every, steps, tops, ncs, X, R, V, W, SL, CO, contacts):
    if cur["step"] % every:
        return
    steps.append(cur["step"]); tops.append(cur["top"]); ncs.append(cur["nc"])
    X.append(cur["x"]); R.append(cur["R"]); V.append(cur["v"]); W.append(cur["w"]); SL.append(cur["sl"]); CO.append(cur["co"])
    contacts.append(cur["c"])


if __name__ == "__main__":
    main()
