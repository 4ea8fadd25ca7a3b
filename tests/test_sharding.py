"""CPU tests of the host-side logic: scene sharding over ranks (world size 2, gloo), batched scene blobs, and that
the C-ABI library loads, exports every symbol include/am3d.h declares and refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adaptivemerging_b200.sharding import reduce_step_stats, shard_scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_scenes_partition():
    for total in (1, 7, 512, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, c = shard_scenes(total, r, world)
                seen += list(range(s, s + c))
            assert seen == list(range(total))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = shard_scenes(4097, rank, world)
    ms, units = reduce_step_stats(10.0 + rank, count * 327, dist)
    q.put((rank, start, count, ms, units))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] + res[1][2] == 4097 and res[1][1] == res[0][2]
    for r in res:
        assert r[3] == 11.0           # MAX over ranks of the timed span
        assert r[4] == 4097 * 327.0   # SUM over ranks of the processed units


def test_replicated_blob_is_consistent():
    from tests.util import golden_scene
    one = golden_scene("tower25platform")
    b = one.replicate(3)
    nb, ns = one.n_bodies, one.n_shapes
    assert b.n_bodies == 3 * nb and b.n_scenes == 3
    assert np.array_equal(b.a["body_scene"], np.repeat(np.arange(3), nb))
    assert np.array_equal(b.a["shape_body"][ns:2 * ns], one.a["shape_body"] + nb)
    assert np.array_equal(b.a["body_shape_first"][2 * nb:], one.a["body_shape_first"] + 2 * ns)
    sp = len(one.a["spring_type"])
    assert np.array_equal(b.a["spring_body1"][sp:2 * sp], one.a["spring_body1"] + nb)


def test_library_exports_and_refuses_cpu():
    from adaptivemerging_b200 import _capi
    L = _capi.load()
    header = open(os.path.join(ROOT, "include", "am3d.h")).read()
    declared = set(re.findall(r"\b(am3d_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), f"libam3d.so does not export {name}"
    assert set(_capi.EXPORTS) <= declared
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert L.am3d_create(0, C.byref(h)) == _capi.ENOGPU  # there is no CPU fallback


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU oracle on a bounded sample, rank 0 only) prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--settle", "5"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "body_steps_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the other ranks of a torchrun launch exit without work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and "{" not in out.stdout
