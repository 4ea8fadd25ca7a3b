#!/usr/bin/env python
"""bench.py — throughput of the AdaptiveMerging 3D rigid-body step on B200.

    python bench.py --gpus N --steps K --warmup W [--workload stack|pile|batch] [--impl reference]

One "step" = one RigidBodySystem.advanceTime(0.05) over the whole workload.  For N > 1 the driver launches
one process per GPU with torch.distributed; every rank steps its own shard of independent scenes (weak
scaling, no data-path collective), the timed region is bracketed by barrier + synchronize and the MAX over
ranks is reported.  Rank 0 prints ONE JSON line.

  value        body-steps/s with the state resident in HBM (device time of K steps, CUDA events on the
               library's own stream)
  e2e          the same metric through the public call with HOST buffers: every step uploads the body
               state from pinned host memory, steps, and downloads the body state
  roofline     PGS sweep kernel: 752 B per contact per iteration / measured sweep time,
               against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (a single-threaded restatement of the reference's Java step) on a bounded
               sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PGS_BYTES_PER_CONTACT_ITER = 752.0
METRIC = "body_steps_per_s"
UNIT = "body-steps/s"


GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_workload(name, size, merging, y0=None):
    """Returns (blob, params, description).  Sizes are per GPU (weak scaling)."""
    from adaptivemerging_b200.ctypes_defs import apply_overrides, default_params
    from adaptivemerging_b200.scene import box_stack, funnel_pile, load_blob
    p = default_params()
    p.enable_merging = int(merging)
    if name == "batch":
        copies = size or 4096
        one = load_blob(os.path.join(GOLDEN, "scene_tower25platform.npz"))
        p = apply_overrides(p, one.overrides)
        blob = one.replicate(copies)
        return blob, p, (f"{copies} independent copies of scenes3D/tower25platform.xml on this GPU (config B of BASELINE.json: "
                         f"4096 copies sharded over the GPUs), merging {'on' if merging else 'off'}")
    if name in ("stack", "pile"):
        n = size or 100
        blob = box_stack(n, n, n, pile=(name == "pile"))
        return blob, p, f"{name} {n}x{n}x{n} unit boxes on a plane (config M of BASELINE.json), merging {'on' if merging else 'off'}"
    if name == "funnel":
        n = size or 20  # BASELINE's config F is n = 100 (100k bodies); see DESIGN.md for why the default is smaller
        tmpl = load_blob(os.path.join(GOLDEN, "scene_funnel_template.npz"))
        p = apply_overrides(p, tmpl.overrides)
        y0 = 0.6 if y0 is None else y0
        blob = funnel_pile(tmpl, nx=n, ny=10, nz=n, y0=y0)
        return blob, p, (f"funnel.xml + {n}x10x{n} torso_flux sphere-tree bodies (config F of BASELINE.json; full size n = 100) with the "
                         f"lowest lattice layer at y0={y0} (SURVEY.md 8d: y0 = 110), merging {'on' if merging else 'off'}")
    raise SystemExit(f"unknown workload {name}")


def sample_workload(name):
    """Bounded CPU sample of the same workload (the reference's broadphase is O(N^2))."""
    from adaptivemerging_b200.scene import box_stack, funnel_pile, load_blob
    if name == "batch":
        return load_blob(os.path.join(GOLDEN, "scene_tower25platform.npz")), "1 copy of tower25platform.xml (328 bodies)"
    if name in ("stack", "pile"):
        return box_stack(12, 100, 12, pile=(name == "pile")), "12x100x12 = 14,400 boxes of the same stack"
    if name == "funnel":
        tmpl = load_blob(os.path.join(GOLDEN, "scene_funnel_template.npz"))
        return funnel_pile(tmpl, nx=10, ny=10, nz=10, y0=0.6), "funnel + 10x10x10 = 1,000 torso bodies piled from y0=0.6"
    raise SystemExit(name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            t = [s.strip() for s in ln.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0]))
                mx.append(float(t[1]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if t[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, workload):
    """DRAM bytes (read + write) per contact per PGS iteration of the sweep kernel, from the committed
    `ncu --set full` capture of the same workload (profiles/r2_traffic.json; dram__bytes_read.sum +
    dram__bytes_write.sum divided by the contact-iterations of the captured launch)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            d = json.load(f)[kernel][{"pile": "stack"}.get(workload, workload)]
        return float(d["dram_bytes_per_contact_iter"]), d["source"]
    except Exception:
        return None, None


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  No JVM exists on the box, so
    this is the single-threaded C++ restatement under oracle/ (kind = "port"), on a bounded sample."""
    if rank != 0:
        return
    from adaptivemerging_b200.ctypes_defs import default_params
    from oracle.oracle import Oracle
    from adaptivemerging_b200.ctypes_defs import apply_overrides
    blob, sample = sample_workload(args.workload)
    p = apply_overrides(default_params(), blob.overrides)
    p.enable_merging = int(args.merging)
    o = Oracle(blob, p)
    for _ in range(args.settle):
        o.step(0.05)
    nb = int((blob.a["body_type"] != 1).sum())
    for _ in range(args.warmup):
        o.step(0.05)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.step(0.05)
    dt = time.perf_counter() - t0
    val = nb * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "sample": sample, "dt": 0.05, "merging": bool(args.merging),
                       "settle_steps": args.settle},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "pgs_row_updates_per_s": o.row_updates() / max(o.solve_seconds(), 1e-12)}
    print(json.dumps(line), flush=True)


def run_leg(local, blob, params, steps, warmup, settle, with_e2e, barrier):
    """One workload on this rank's GPU: settle + warm-up, K timed steps with the state resident in HBM, then (with_e2e)
    reset, the same settle + warm-up again and the SAME K steps timed end to end through host buffers."""
    import torch
    from adaptivemerging_b200.system import RigidBodySystem
    sysm = RigidBodySystem(local).load(blob, params)
    sysm.set_option("record_events", 0)  # the merge / unmerge event log is a parity-test aid
    for kv in filter(None, os.environ.get("AM3D_OPTIONS", "").split(",")):  # A/B runs: AM3D_OPTIONS=pgs_fast_rows=0,...
        sysm.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    for _ in range(settle + warmup):
        sysm.advanceTime(0.05)
    s0 = sysm.stats()
    profiling = bool(os.environ.get("AM3D_CUDA_PROFILER"))  # ncu --profile-from-start off: capture the timed region only
    if profiling:
        torch.cuda.cudart().cudaProfilerStart()
    clocks = ClockSampler(local)
    clocks.start()
    barrier()
    sysm.mark(0)
    t0 = time.perf_counter()
    np_bytes = np_s = 0.0
    sum_contacts = 0
    series = []
    for _ in range(steps):
        sysm.advanceTime(0.05)
        t = sysm.timings()
        sum_contacts += t.n_contacts
        series.append(round(t.compute_time * 1e3, 2))
        # narrowphase, SURVEY.md 8d: 8 B + 2 x 96 B per candidate pair, 80 B per emitted contact
        np_bytes += 200.0 * t.n_pairs + 80.0 * t.n_contacts
        np_s += t.narrowphase_kernel_time
    sysm.mark(1)
    ms = sysm.elapsed_ms()
    barrier()
    wall = time.perf_counter() - t0
    if profiling:
        torch.cuda.cudart().cudaProfilerStop()
    clk = clocks.stop()
    s1 = sysm.stats()
    tm = sysm.timings()
    out = {"ms": ms, "wall": wall, "clocks": clk, "np_bytes": np_bytes, "np_s": np_s, "sum_contacts": sum_contacts,
           "row_updates": s1["row_updates"] - s0["row_updates"], "solve_s": s1["solve_seconds"] - s0["solve_seconds"],
           "series": series, "launches": s1["kernel_launches"] - s0["kernel_launches"], "solve_launches": s1["solve_launches"] - s0["solve_launches"],
           "tm": {"n_collections": tm.n_collections, "n_contacts": tm.n_contacts, "n_pairs": tm.n_pairs, "pgs_colors": tm.pgs_colors,
                  "n_bodies_top_level": tm.n_bodies, "pgs_kernel": tm.pgs_kernel, "pgs_giant_groups": tm.pgs_giant_groups,
                  "phase_ms": {"detection": tm.detection * 1e3, "warmstart": tm.warmstart * 1e3, "update_collections": tm.update_collections * 1e3,
                               "contact_ordering": tm.contact_ordering * 1e3, "single_it_pgs": tm.single_it_pgs * 1e3,
                               "unmerging": tm.unmerging * 1e3, "lcp_solve": tm.lcp_solve * 1e3, "pgs_sweeps": tm.pgs_kernel_time * 1e3,
                               "merging": tm.merging * 1e3, "total": tm.compute_time * 1e3}},
           "ms_e2e": None, "h2d": 0, "d2h": 0}
    if with_e2e:
        # ---- end-to-end leg: host buffers in, host buffers out, every step, over the SAME simulation steps ----------
        # inputs of a step as the Java front end hands them over: per-body velocity pokes (mouse impulses / scripted
        # pushes; zeros here) from pinned host memory; result: the full body state for drawing
        sysm.reset()
        sysm.set_option("record_events", 0)
        for _ in range(settle + warmup):
            sysm.advanceTime(0.05)
        n = sysm.n_bodies
        poke_v = torch.zeros((n, 3), dtype=torch.float64).pin_memory()
        poke_w = torch.zeros((n, 3), dtype=torch.float64).pin_memory()
        pv, pw = poke_v.numpy(), poke_w.numpy()
        out["h2d"] = pv.nbytes + pw.nbytes
        out["d2h"] = n * (3 + 9 + 3 + 3) * 8 + 2 * 4 * n

        def pinned_state():
            t = {"x": torch.empty((n, 3), dtype=torch.float64).pin_memory(), "R": torch.empty((n, 9), dtype=torch.float64).pin_memory(),
                 "v": torch.empty((n, 3), dtype=torch.float64).pin_memory(), "omega": torch.empty((n, 3), dtype=torch.float64).pin_memory(),
                 "sleeping": torch.empty(n, dtype=torch.int32).pin_memory(), "collection": torch.empty(n, dtype=torch.int32).pin_memory()}
            return t, {k: a.numpy() for k, a in t.items()}
        # two result buffers: the copy of step N's state (second stream) runs under the kernels of step N+1, as a front end
        # that draws one frame behind would use it; every step's state is fully delivered inside the timed region
        keep, states = zip(*(pinned_state() for _ in range(2)))
        barrier()
        sysm.mark(0)
        for k in range(steps):
            sysm.add_velocities(pv, pw)
            sysm.advanceTime(0.05)
            sysm.bodies_async(states[k & 1])
        sysm.wait_bodies()
        sysm.mark(1)
        out["ms_e2e"] = sysm.elapsed_ms()
        barrier()
        out["e2e_contacts_last_step"] = sysm.timings().n_contacts
    sysm.close()
    return out


def roofline_of(leg, workload, steps):
    peak, peak_src = hbm_peak()
    row_updates, solve_s = leg["row_updates"], leg["solve_s"]
    achieved = (row_updates / 3.0) * PGS_BYTES_PER_CONTACT_ITER / max(solve_s, 1e-12) / 1e9
    solve_launches = max(float(leg["solve_launches"]), 1.0)
    kernel = {0: "k_pgs_color<1>", 1: "k_pgs_persistent", 2: "k_pgs_cluster"}[leg["tm"]["pgs_kernel"]]
    if leg["tm"]["pgs_giant_groups"]:
        kernel = "k_pgs_giant<1> + " + kernel
    traffic_ratio, traffic_src = measured_traffic(kernel, workload)
    contact_iters_per_launch = (row_updates / 3.0) / solve_launches
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None if traffic_ratio is None else traffic_ratio * contact_iters_per_launch,
            "traffic_source": traffic_src, "peak_source": peak_src,
            "algorithmic_bytes": "752 B per contact per PGS iteration (SURVEY.md 8d)",
            "algorithmic_bytes_per_launch": PGS_BYTES_PER_CONTACT_ITER * contact_iters_per_launch,
            "launches": int(solve_launches), "avg_launch_ms": 1e3 * solve_s / solve_launches,
            "sweep_ms_per_step": 1e3 * solve_s / steps}


def narrow_roofline_of(leg, steps):
    peak, _ = hbm_peak()
    a = leg["np_bytes"] / max(leg["np_s"], 1e-12) / 1e9
    return {"bound": "hbm", "kernel": "k_narrow_box + k_narrow_tree<0/1>", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak,
            "algorithmic_bytes": "200 B per candidate pair + 80 B per contact (SURVEY.md 8d)", "ms_per_step": 1e3 * leg["np_s"] / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="batch", choices=["batch", "stack", "pile", "funnel"])
    ap.add_argument("--merging", type=int, default=1)
    ap.add_argument("--size", type=int, default=0, help="per-GPU size: copies (batch), edge length (stack / pile), lattice edge (funnel)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="batch only: strong = BASELINE.json's config B as written, 4096 copies in total, 4096/N per GPU; "
                         "weak = 512 copies per GPU")
    ap.add_argument("--y0", type=float, default=None, help="funnel: height of the lowest lattice layer (SURVEY.md: 110)")
    ap.add_argument("--settle", type=int, default=-1,
                    help="untimed steps before the warm-up so that the workload is in its loaded phase "
                         "(default: 120 batch = towers collapsing onto the platform with the first collections formed, 20 stack/pile, 35 funnel = second layer "
                         "landing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the short config M / config F legs reported under `also`")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.settle < 0:
        args.settle = {"batch": 120, "stack": 20, "pile": 20, "funnel": 35}[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the rigid-body step has no CPU fallback")

    from adaptivemerging_b200.sharding import reduce_stats, shard_scenes
    size = args.size
    if args.workload == "batch" and not size:
        # config B: 4096 independent scenes, a contiguous block of scene ids per GPU (no data-path collective)
        size = shard_scenes(4096, rank, world)[1] if args.scaling == "strong" else 512
    blob, params, desc = build_workload(args.workload, size, args.merging, args.y0)
    nb = int((blob.a["body_type"] != 1).sum())  # non-plane leaf bodies

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    leg = run_leg(local, blob, params, args.steps, args.warmup, args.settle, True, barrier)
    del blob

    # MAX over ranks of the timed spans, SUM over ranks of the processed units (adaptivemerging_b200/sharding.py)
    counts = [float(nb), float(leg["row_updates"]), float(leg["solve_s"]), float(leg["launches"])]
    times, tot = reduce_stats([leg["ms"], leg["ms_e2e"]], counts, dist, f"cuda:{local}")
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    ms, ms_e2e = float(times[0]), float(times[1])
    total_bodies = float(tot[0])
    value = total_bodies * args.steps / (ms * 1e-3)
    e2e = total_bodies * args.steps / (ms_e2e * 1e-3)
    tm = leg["tm"]
    rec_mb = tm["n_contacts"] * 192.0 / 1e6
    scaling = args.scaling if args.workload == "batch" and not args.size else "weak"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "bodies_per_gpu": nb, "bodies_total": int(total_bodies), "dt": 0.05,
                   "merging": bool(args.merging), "pgs_iterations": params.iterations, "settle_steps": args.settle,
                   "l2": f"the PGS contact records alone are {rec_mb:.0f} MB per sweep ({'larger than' if rec_mb > 126 else 'within'} the 126 MB "
                         "L2) and every step re-detects its contacts; no flush between steps (steps are data dependent)",
                   "e2e_window": "the end-to-end leg repeats the resident leg's simulation steps (reset, same settle + warm-up)"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": leg["h2d"], "d2h_bytes_per_step": leg["d2h"],
                "ms_per_step": ms_e2e / args.steps, "contacts_last_step": leg.get("e2e_contacts_last_step")},
        "gpu_launches": int(float(tot[3])),
        "clocks": leg["clocks"],
        "pgs_row_updates_per_s": float(tot[1]) / max(float(counts[2]), 1e-12),
        # BASELINE.md row "PGS contact-row updates/s": 23.4 M/s (authors' Java log, tower25platform, 30 iterations, unknown CPU);
        # body-steps/s has no published number for leaf bodies, so vs_baseline stays null
        "pgs_row_updates_vs_baseline": (float(tot[1]) / max(float(counts[2]), 1e-12)) / 23.4e6,
        "collections_last_step": tm["n_collections"], "contacts_last_step": tm["n_contacts"], "pairs_last_step": tm["n_pairs"],
        "pgs_phases": tm["pgs_colors"], "phase_ms_last_step": tm["phase_ms"],
        "wall_ms_per_step": 1e3 * leg["wall"] / args.steps, "step_ms_series": leg["series"],
        "roofline": roofline_of(leg, args.workload, args.steps),
        "roofline_narrowphase": narrow_roofline_of(leg, args.steps),
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle.oracle import Oracle
        sblob, sample = sample_workload(args.workload)
        o = Oracle(sblob, params)
        snb = int((sblob.a["body_type"] != 1).sum())
        for _ in range(args.settle):
            o.step(0.05)
        t0 = time.perf_counter()
        k = 0
        while time.perf_counter() - t0 < 15.0 and k < 50:
            o.step(0.05)
            k += 1
        el = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": snb * k / el, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{sample}, {k} steps after {args.settle} settle steps",
                                "pgs_row_updates_per_s": o.row_updates() / max(o.solve_seconds(), 1e-12)}
    if world == 1 and not args.no_also and args.workload == "batch" and not args.size:
        # short legs of BASELINE.json's other single-GPU configs, so that one default run shows them all
        also = []
        for wl, sz, mg, st, settle, y0 in [("stack", 100, 0, 5, 20, None), ("stack", 100, 1, 5, 20, None), ("funnel", 20, 1, 3, 35, 0.6)]:
            try:
                b2, p2, d2 = build_workload(wl, sz, mg, y0)
                nb2 = int((b2.a["body_type"] != 1).sum())
                l2 = run_leg(local, b2, p2, st, 3, settle, False, barrier)
                del b2
                also.append({"workload": wl, "description": d2, "bodies": nb2, "merging": bool(mg), "steps": st, "settle_steps": settle,
                             "value": nb2 * st / (l2["ms"] * 1e-3), "unit": UNIT, "ms_per_step": l2["ms"] / st,
                             "pgs_row_updates_per_s": l2["row_updates"] / max(l2["solve_s"], 1e-12),
                             "contacts_last_step": l2["tm"]["n_contacts"], "collections_last_step": l2["tm"]["n_collections"],
                             "roofline": roofline_of(l2, wl, st), "roofline_narrowphase": narrow_roofline_of(l2, st)})
            except Exception as e:  # an auxiliary leg must not take the headline line down
                also.append({"workload": wl, "error": str(e)[:200]})
        line["also"] = also
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
