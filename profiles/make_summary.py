#!/usr/bin/env python3
"""Rebuild profiles/r1_final_ncu_summary.md from the ncu exports in gpurun_out/ (launch list csv + `--page raw --csv`
of the full capture) and the bench JSON lines copied into profiles/.  Usage: python profiles/make_summary.py"""
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = lambda *p: os.path.join(ROOT, "gpurun_out", *p)
P = lambda *p: os.path.join(ROOT, "profiles", *p)
out = []
out.append("# Round 1 - ncu evidence, final state of the round\n")
out.append("Command (1 GPU, B200, `--clock-control none`): `python bench.py --steps 1 --warmup 1 --scale 0.25 --no-cpu-baseline`")
out.append("(C3 at 1/4 scale: 50 000 transcripts, 112.7 M nt, 2.48 M molecules, 2.5 M requested fragments; 2 end-to-end + 2 device-resident passes).")
out.append("The same command exited 0 without ncu first. Per-launch times under ncu are cold-cache and serialised: compare shares.\n")
rows = [r for r in csv.reader(open(P("r1_final_launches_ncu.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    short = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[ki])) or "(memset/memcpy)"
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
out.append("## Launch list (profiles/r1_final_launches_ncu.csv): %d launches\n" % (len(rows) - 1))
out.append("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    out.append("| `%s` | %d | %.1f | %.1f%% |" % (k, c, t, 100 * t / tot))
d = json.load(open(P("r1_final_bench.json")))
out.append("\nCUDA-event stage times of the un-profiled full-scale run (profiles/r1_final_bench.json), ms per step of %.2f ms:" % d["ms_per_step"])
out.append(", ".join("%s %.2f" % (k, v) for k, v in sorted(d["kernels_ms_per_step"].items(), key=lambda kv: -kv[1])) + ".")
out.append("Under ncu `k_prefix` / `k_ingest` appear once per 32 MB chunk in the end-to-end passes (there they overlap the H2D copy) and `k_prefix` once per device-resident pass.\n")
rows = list(csv.reader(open(G("prof_r1_final_raw.csv"))))
hdr, units = rows[0], rows[1]
g = lambda r, n: r[hdr.index(n)]
u = lambda n: units[hdr.index(n)]
st = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
best = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", g(r, "Kernel Name"))
    if name not in best or float(g(best[name], "gpu__time_duration.sum")) < float(g(r, "gpu__time_duration.sum")):
        best[name] = r
out.append("## `ncu --set full` (gpurun_out/prof_r1_final.ncu-rep), largest launch of each kernel, scale 0.25\n")
out.append("| kernel | time %s | DRAM read %s | DRAM write %s | DRAM %% peak | issue active %% | FP64 pipe %% | regs | warps active %% | active lanes / inst | top stalls |"
           % (u("gpu__time_duration.sum"), u("dram__bytes_read.sum"), u("dram__bytes_write.sum")))
out.append("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for name, r in sorted(best.items(), key=lambda kv: -float(g(kv[1], "gpu__time_duration.sum"))):
    t = sum(float(g(r, h).replace(",", "") or 0) for h in st)
    top = sorted(((float(g(r, h).replace(",", "") or 0), h) for h in st), reverse=True)[:3]
    tops = ", ".join("%s %.0f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / t) for v, h in top)
    out.append("| `%s` | %.1f | %.3f | %.3f | %.1f | %.1f | %.1f | %s | %.1f | %.1f | %s |" % (
        name, float(g(r, "gpu__time_duration.sum")), float(g(r, "dram__bytes_read.sum")), float(g(r, "dram__bytes_write.sum")),
        float(g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")), float(g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")),
        float(g(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")), g(r, "launch__registers_per_thread"),
        float(g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")), float(g(r, "smsp__thread_inst_executed_per_inst_executed.ratio")), tops))
n2, n4, ref = (json.load(open(P(f))) for f in ("r1_final_bench_n2.json", "r1_final_bench_n4.json", "r1_final_bench_reference.json"))
out.append("""
## What changed during the round (device-resident C3 step, one B200; CUDA-event stage times in ms)

| state | step | k_fragment + k_intervals | k_prefix | k_pcr | note |
|---|---:|---:|---:|---:|---|
| first correct path (profiles/r1_ncu_summary.md) | 21.6 | 9.04 | 4.89 | 4.32 | thread per molecule does everything; CTA per transcript reloads the LUT; per-cycle binomial |
| + interval work queue (`k_fragment` -> `k_intervals`) | 18.1 | 3.82 | 5.34 | 5.59 | latency-bound searches run lean, at high occupancy |
| + persistent prefix CTAs, warp-synchronous PCR | 15.3 | 3.73 | 4.15 | 4.03 | |
| + warp-per-transcript prefix, per-warp transpose for 16 B stores | 13.4 | 3.74 | 2.04 | 4.03 | |
| + PCR in warp rounds with batched exact tests, interpolation search for priming | 11.2 | 2.93 | 2.05 | 2.56 | |
| + two-level prefix sums (u32 in-block + u64 per block), constant-shift fast path | 9.7 | 2.93 | 1.15 | 2.57 | half the prefix bytes |
| + prefetched header draws, occupancy-targeted launch bounds | 8.6 | 2.17 | 1.15 | 2.26 | |
| + FP32 pre-decision of the BTRS acceptance test (exact FP64 fallback) | %.1f | %.2f | %.2f | %.2f | final |

End to end (host buffers in, fragment records out): 43.8 ms -> %.1f ms (persistent staging, H2D on a copy stream overlapped with
ingest + prefix sums per 32 MB chunk, global transcript index written on the device).
Weak scaling (profiles/r1_final_bench_n2.json, _n4.json): %.2f / %.2f / %.2f G fragments/s at 1 / 2 / 4 GPUs (%.0f %% / %.0f %% of linear).
Reference release binary on the same box, %d cores (profiles/r1_final_bench_reference.json): %.2e fragments/s.""" % (
    d["ms_per_step"], d["kernels_ms_per_step"]["fragment"] + d["kernels_ms_per_step"]["intervals"], d["kernels_ms_per_step"]["prefix"],
    d["kernels_ms_per_step"]["pcr"], d["e2e"]["ms_per_step"], d["value"] / 1e9, n2["value"] / 1e9, n4["value"] / 1e9,
    100 * n2["value"] / (2 * d["value"]), 100 * n4["value"] / (4 * d["value"]), ref["cpu_baseline"]["cores"], ref["value"]))
open(P("r1_final_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-40:]))
