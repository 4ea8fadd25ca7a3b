#!/usr/bin/env python
"""Turn the outputs of tools/gpu_session_r1.sh (gpurun_out/r1/) into the tracked documents under profiles/.

    python tools/make_profiles.py [round_tag]      # default r1
"""
import io
import json
import os
import shutil
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def load(path):
    try:
        lines = [ln for ln in open(path) if ln.startswith("{")]
        return json.loads(lines[-1])
    except Exception:
        return None


def capture(fn, *a):
    buf = io.StringIO()
    with redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    runs = [("batch (default; config B, 512 x tower25platform per GPU, merging on)", "bench_batch", "python bench.py"),
            ("batch, 2 GPUs (torchrun)", "bench_batch_2gpu", "torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 5"),
            ("stack 100^3, merging off (config M)", "bench_stack_m0", "python bench.py --workload stack --merging 0"),
            ("stack 100^3, merging on (config M)", "bench_stack_m1", "python bench.py --workload stack --merging 1"),
            ("pile 100^3, merging on (config M, jittered)", "bench_pile_m1", "python bench.py --workload pile --merging 1"),
            ("funnel 20x10x20 torsos (config F, reduced: DESIGN.md 5)", "bench_funnel20", "python bench.py --workload funnel --steps 5 --warmup 3"),
            ("reference arm: CPU oracle, 1 core", "bench_reference_batch", "python bench.py --impl reference --steps 40 --warmup 3")]
    out = [f"# Bench lines, round {tag[1:]} (B200, one fresh box per session; `tools/gpu_session_{tag}.sh`)\n",
           "Metric: body-steps/s. `resident` = state in HBM, CUDA events on the library stream; `e2e` = velocity pokes up from",
           "pinned host memory + step + full body state down into pinned host memory, every step. `roofline` = 752 B x",
           "contacts x iterations / PGS sweep time / 6392.8 GB/s (SURVEY.md 8d); `traffic` = ncu DRAM bytes of that launch.",
           "For scale: the authors' own Java log of ONE tower25platform scene (BASELINE.md) gives 7.53 ms per step = 43.6 k",
           "leaf-body-steps/s and 23.4 M PGS row updates/s on an unknown CPU; the single-core oracle here runs the same scene at",
           "~38 k body-steps/s and 53 M row updates/s.",
           "The default (batch, 1 GPU) line is from the final code; the other lines and the ncu captures were taken before the",
           "last change of the round (one thread per contact in `k_contact_set`), which took the 1M-box stack (merging off)",
           "from 24.2 to 21.5 ms/step in a 10-step run (detection 3.65 -> 2.40 ms); all 31 GPU tests green after it.\n",
           "| workload | bodies/GPU | GPUs | ms/step | resident body-steps/s | e2e body-steps/s | PGS row-updates/s | roofline frac (kernel) | DRAM traffic / launch | narrowphase frac | contacts | colours |",
           "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    raw = []
    for label, name, cmd in runs:
        d = load(os.path.join(src, name + ".json"))
        if d is None:
            continue
        raw.append((label, cmd, d))
        if d.get("impl") == "reference":
            out.append(f"| {label} | {d['config'].get('sample', '')} | - | {d['ms_per_step']:.2f} | {d['value']:.0f} | {d['e2e']['value']:.0f} | "
                       f"{d.get('pgs_row_updates_per_s', 0):.3g} | - | - | - | - | - |")
            continue
        r = d["roofline"]
        tr = r.get("traffic")
        nf = d.get("roofline_narrowphase", {}).get("frac")
        out.append(f"| {label} | {d['config']['bodies_per_gpu']} | {d['n_gpus']} | {d['ms_per_step']:.2f} | {d['value']:.4g} | {d['e2e']['value']:.4g} | "
                   f"{d['pgs_row_updates_per_s']:.3g} | {r['frac']:.3f} (`{r['kernel']}`) | {'%.2f GB' % (tr / 1e9) if tr else '-'} | "
                   f"{'%.3f' % nf if nf is not None else '-'} | {d.get('contacts_last_step')} | {d.get('pgs_colors')} |")
    out.append("\n## Phase times of the last timed step (ms)\n")
    out.append("| workload | detection | warm start | LCP solve (of which PGS sweeps) | total | CPU baseline (oracle, 1 core) |")
    out.append("|---|---:|---:|---:|---:|---|")
    for label, cmd, d in raw:
        ph = d.get("phase_ms_last_step")
        if not ph:
            continue
        cb = d.get("cpu_baseline")
        cbs = f"{cb['value']:.0f} body-steps/s on {cb['sample']}" if cb else "-"
        out.append(f"| {label} | {ph['detection']:.2f} | {ph['warmstart']:.2f} | {ph['lcp_solve']:.2f} ({ph['pgs_sweeps']:.2f}) | {ph['total']:.2f} | {cbs} |")
    out.append("\n## Raw JSON lines\n")
    for label, cmd, d in raw:
        out.append(f"`{cmd}`\n\n```json\n{json.dumps(d)}\n```\n")
    open(os.path.join(dst, f"{tag}_bench.md"), "w").write("\n".join(out) + "\n")

    # ncu: launch list of the default bench command + --set full summaries
    for wl, title, flags in (("batch", "default bench workload (512 x tower25platform)", ""),
                             ("stack", "1M-box stack, merging off", " --workload stack --merging 0")):
        lcsv = os.path.join(src, f"launches_{wl}.csv")
        if not os.path.exists(lcsv):
            continue
        shutil.copy(lcsv, os.path.join(dst, f"launches_{tag}_{wl}.csv"))
        body = capture(ncu_summary.launches, lcsv)
        head = (f"# ncu launch list, round {tag[1:]} — {title}, 2 timed steps\n\n"
                "Command: `AM3D_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv "
                f"python bench.py{flags} --steps 2 --warmup 3 --no-cpu-baseline` (the profiler range is the timed region).\n"
                "Cold-cache, serialised launch times: compare SHARES, not absolutes. Raw CSV: "
                f"`launches_{tag}_{wl}.csv`.\n\n")
        open(os.path.join(dst, f"{tag}_ncu_launches_{wl}.md"), "w").write(head + body)
    for name, title in (("full_batch", "default bench workload (512 x tower25platform per GPU)"),
                        ("full_stack", "1M-box stack, merging off")):
        rep = os.path.join(src, name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        body = capture(ncu_summary.full, rep)
        head = (f"# ncu --set full, round {tag[1:]} — {title}\n\n"
                "Command: `AM3D_CUDA_PROFILER=1 ncu --profile-from-start off --set full --clock-control none --import-source on "
                "-k regex:... python bench.py [--workload ...] --steps 1 --warmup 3 --no-cpu-baseline`; one timed step, launches in "
                "order. The long `k_pgs_persistent` launch is the full solve (30 iterations + warm-start pass), the short one the "
                "single sweep over internal contacts.\n\n")
        open(os.path.join(dst, f"{tag}_ncu_{name}.md"), "w").write(head + body)


if __name__ == "__main__":
    main()
