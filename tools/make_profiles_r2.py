#!/usr/bin/env python
"""Turn the outputs of tools/gpu_session_r2.sh (gpurun_out/r2/) into the tracked documents under profiles/:
r2_bench.md (bench lines), r2_ncu_launches_*.md (+ raw CSV), r2_ncu_full_*.md, r2_traffic.json, r2_funnel50.jsonl."""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

SRC = os.path.join(ROOT, "gpurun_out", "r2")
DST = os.path.join(ROOT, "profiles")


def load(name):
    try:
        lines = [ln for ln in open(os.path.join(SRC, name)) if ln.startswith("{")]
        return json.loads(lines[-1])
    except Exception:
        return None


def capture(fn, *a):
    buf = io.StringIO()
    with redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


def raw_rows(rep):
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return [], {}
    ix = {h: i for i, h in enumerate(rows[0])}
    raw_rows.units = {h: rows[1][i] for h, i in ix.items()}
    return rows[2:], ix


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    os.makedirs(DST, exist_ok=True)
    runs = [("B: 4096 x tower25platform on one GPU (default; config B as written)", "bench_batch.json", "python bench.py"),
            ("B, 512 copies per GPU (--scaling weak; round 1's workload)", "bench_batch512.json", "python bench.py --scaling weak --no-also --no-cpu-baseline"),
            ("M: 1 M-box stack, merging off", "bench_stack_m0.json", "python bench.py --workload stack --merging 0 --no-cpu-baseline"),
            ("M: 1 M-box stack, merging on", "bench_stack_m1.json", "python bench.py --workload stack --merging 1 --no-cpu-baseline"),
            ("M: 1 M-box pile (jittered), merging on", "bench_pile_m1.json", "python bench.py --workload pile --merging 1 --no-cpu-baseline"),
            ("F: funnel + 4 000 torsos (20x10x20), y0 = 0.6", "bench_funnel20.json", "python bench.py --workload funnel --steps 5 --warmup 3 --no-cpu-baseline"),
            ("reference arm: CPU oracle, 1 core", "bench_reference_batch.json", "python bench.py --impl reference --steps 40 --warmup 3")]
    out = ["# Bench lines, round 2 (B200, one fresh box per session; `tools/gpu_session_r2.sh`)\n",
           "Metric: body-steps/s. `resident` = state in HBM, CUDA events on the library stream; `e2e` = the same simulation steps",
           "with velocity pokes up from pinned host memory and the full body state down into pinned host memory, every step.",
           "`roofline frac` = 752 B x contact-iterations / PGS sweep time / the measured HBM peak of MEASURED_PEAKS.json (SURVEY.md 8d); `traffic` = ncu DRAM bytes",
           "per contact-iteration (r2_traffic.json) x the contact-iterations of a launch.\n"]
    for f in ("pytest_gpu.txt", "smoke.txt"):
        pth = os.path.join(SRC, f)
        if os.path.exists(pth):
            out.append(f"`{f}`: " + " | ".join(ln.strip() for ln in open(pth) if ln.strip()) + "\n")
    out += ["| workload | bodies/GPU | GPUs | ms/step | resident body-steps/s | e2e body-steps/s | PGS row-updates/s | roofline frac (kernel) | sweep ms/step | narrowphase frac | contacts | collections | launches/step |",
            "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    raw = []
    for label, name, cmd in runs:
        d = load(name)
        if d is None:
            continue
        raw.append((label, cmd, d))
        if d.get("impl") == "reference":
            out.append(f"| {label} | {d['config'].get('sample', '')} | - | {d['ms_per_step']:.2f} | {d['value']:.0f} | {d['e2e']['value']:.0f} | "
                       f"{d.get('pgs_row_updates_per_s', 0):.3g} | - | - | - | - | - | - |")
            continue
        r = d["roofline"]
        nf = d.get("roofline_narrowphase", {}).get("frac")
        out.append(f"| {label} | {d['config']['bodies_per_gpu']} | {d['n_gpus']} | {d['ms_per_step']:.2f} | {d['value']:.4g} | {d['e2e']['value']:.4g} | "
                   f"{d['pgs_row_updates_per_s']:.3g} | {r['frac']:.3f} (`{r['kernel']}`) | {r.get('sweep_ms_per_step', 0):.2f} | "
                   f"{'%.3f' % nf if nf is not None else '-'} | {d.get('contacts_last_step')} | {d.get('collections_last_step')} | {d['gpu_launches'] / d['steps']:.0f} |")
    d = load("bench_batch.json")
    if d and d.get("also"):
        out.append("\n## Short legs inside the default line (`also`)\n")
        out.append("| workload | bodies | merging | ms/step | body-steps/s | PGS row-updates/s | roofline frac | narrowphase frac | contacts |")
        out.append("|---|---:|---|---:|---:|---:|---:|---:|---:|")
        for a in d["also"]:
            if "error" in a and a.get("error"):
                out.append(f"| {a['workload']} | error: {a['error']} |")
                continue
            out.append(f"| {a['workload']} | {a['bodies']} | {a['merging']} | {a['ms_per_step']:.2f} | {a['value']:.4g} | {a['pgs_row_updates_per_s']:.3g} | "
                       f"{a['roofline']['frac']:.3f} | {a['roofline_narrowphase']['frac']:.3f} | {a['contacts_last_step']} |")
    out.append("\n## Phase times of the last timed step (ms)\n")
    for label, cmd, d in raw:
        ph = d.get("phase_ms_last_step")
        if ph:
            out.append(f"* {label}: " + ", ".join(f"{k} {v:.2f}" for k, v in ph.items()))
    out.append("\n## Raw JSON lines\n")
    for label, cmd, d in raw:
        out.append(f"`{cmd}`\n\n```json\n{json.dumps(d)}\n```\n")
    open(os.path.join(DST, "r2_bench.md"), "w").write("\n".join(out) + "\n")

    for wl, title, flags in (("batch", "default bench workload (4096 x tower25platform)", " --no-also"),
                             ("batch512", "512 x tower25platform (--scaling weak)", " --scaling weak --no-also"),
                             ("stack", "1M-box stack, merging off", " --workload stack --merging 0"),
                             ("funnel_tree", "funnel + 4 000 torsos: the sphere-tree narrowphase kernels only (-k regex:\"k_narrow|k_tree\", 1 step)",
                              " --workload funnel")):
        lcsv = os.path.join(SRC, f"launches_{wl}.csv")
        if not os.path.exists(lcsv):
            continue
        shutil.copy(lcsv, os.path.join(DST, f"launches_r2_{wl}.csv"))
        body = capture(ncu_summary.launches, lcsv)
        head = (f"# ncu launch list, round 2 — {title}, 2 timed steps\n\n"
                "Command: `AM3D_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv "
                f"python bench.py{flags} --steps 2 --warmup 3 --no-cpu-baseline` (the profiler range is the timed region).\n"
                f"Cold-cache, serialised launch times: compare SHARES, not absolutes. Raw CSV: `launches_r2_{wl}.csv`.\n\n")
        open(os.path.join(DST, f"r2_ncu_launches_{wl}.md"), "w").write(head + body)
    traffic = {}
    for name, title, wl, bench in (("full_batch", "default bench workload (4096 x tower25platform)", "batch", "bench_batch.json"),
                                   ("full_batch512", "512 x tower25platform (--scaling weak)", "batch", "bench_batch512.json"),
                                   ("full_stack", "1M-box stack, merging off", "stack", "bench_stack_m0.json"),
                                   ("full_funnel", "funnel + 4 000 torsos", "funnel", "bench_funnel20.json")):
        rep = os.path.join(SRC, name + "_raw.csv")   # raw page exported on the GPU box
        if not os.path.exists(rep):
            rep = os.path.join(SRC, name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        body = capture(ncu_summary.full, rep)
        # keep the summaries readable: at most the first 3 launches of every kernel
        seen, kept, block = {}, [], []
        for ln in body.splitlines(keepends=True):
            if ln.startswith("### "):
                if block:
                    kept.append("".join(block))
                block = [ln]
                k = ln
                seen[k] = seen.get(k, 0) + 1
                if seen[k] > 3:
                    block = None
            elif block is not None:
                block.append(ln)
        if block:
            kept.append("".join(block))
        head = (f"# ncu --set full, round 2 — {title}\n\n"
                "Command: `AM3D_CUDA_PROFILER=1 ncu --profile-from-start off --set full --clock-control none --import-source on "
                "-k regex:... python bench.py [...] --steps 1 --warmup 3 --no-cpu-baseline`; one timed step, launches in order "
                "(at most three launches per kernel shown).\n\n")
        open(os.path.join(DST, f"r2_ncu_{name}.md"), "w").write(head + "".join(kept))
        # DRAM traffic of the sweep kernels per contact-iteration
        rows, ix = raw_rows(rep)
        bd = load(bench)
        if not rows or bd is None:
            continue
        contacts = bd["contacts_last_step"]
        try:  # the profiled run prints its own line: use ITS contact count
            lg = [ln for ln in open(os.path.join(SRC, "ncu_" + name + ".log")) if ln.startswith("{")]
            contacts = json.loads(lg[-1])["contacts_last_step"]
        except Exception:
            pass
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        ur = scale.get(raw_rows.units.get("dram__bytes_read.sum", ""), None)
        uw = scale.get(raw_rows.units.get("dram__bytes_write.sum", ""), None)
        for kern, key in (("k_pgs_color<1", "k_pgs_color<1>"), ("k_pgs_persistent", "k_pgs_persistent")):
            sel = [r for r in rows if kern in r[ix["Kernel Name"]]]
            if not sel:
                continue
            if key == "k_pgs_color<1>":
                # ONE iteration: the per-phase launches up to (and with) the launch that folds the trailing phases
                # (k_pgs_tail<1>), or `pgs_phases` launches when there is none
                seq = [r for r in rows if "k_pgs_color<1" in r[ix["Kernel Name"]] or "k_pgs_tail<1" in r[ix["Kernel Name"]]]
                cut = next((i for i, r in enumerate(seq) if "k_pgs_tail<1" in r[ix["Kernel Name"]]), None)
                if cut is not None:
                    sel = seq[:cut + 1]
                else:
                    phases = bd.get("pgs_phases", len(sel))
                    sel = sel[:phases]
                    if len(sel) < phases:
                        continue
                iters = 1
            else:
                sel = [max(sel, key=lambda r: fnum(r[ix["gpu__time_duration.sum"]]))]
                iters = 30
            rd = sum(fnum(r[ix["dram__bytes_read.sum"]]) for r in sel)
            wr = sum(fnum(r[ix["dram__bytes_write.sum"]]) for r in sel)
            ms = sum(fnum(r[ix["gpu__time_duration.sum"]]) for r in sel)
            traffic.setdefault(key, {})[wl if name != "full_batch512" else "batch512"] = {
                "dram_bytes_per_contact_iter": None if ur is None or uw is None else (rd * ur + wr * uw) / (contacts * iters),
                "raw_read": rd, "raw_write": wr, "raw_time": ms, "launches": len(sel),
                "contacts": contacts, "iterations": iters,
                "source": f"ncu --set full, profiles/r2_ncu_{name}.md: DRAM read + write of {len(sel)} launch(es) "
                          f"({raw_rows.units.get('dram__bytes_read.sum')}) / ({contacts} contacts x {iters} iteration(s))"}
    if traffic:
        old = {}
        if os.path.exists(os.path.join(DST, "r2_traffic.json")):  # keep the entries of earlier sessions' captures
            old = json.load(open(os.path.join(DST, "r2_traffic.json")))
        for k, v in traffic.items():
            old.setdefault(k, {}).update(v)
        json.dump(old, open(os.path.join(DST, "r2_traffic.json"), "w"), indent=1)
    f50 = os.path.join(SRC, "funnel50.jsonl")
    if os.path.exists(f50):
        shutil.copy(f50, os.path.join(DST, "r2_funnel50.jsonl"))
    print("profiles written:", sorted(f for f in os.listdir(DST) if f.startswith("r2_") or "_r2_" in f))


if __name__ == "__main__":
    main()
