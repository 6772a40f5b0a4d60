#!/usr/bin/env python3
"""Rebuild profiles/r2_ncu_summary.md and profiles/ncu_traffic.json from the ncu exports brought back in gpurun_out/
by `bash profiles/r2_final_commands.sh` (the exact commands of the final GPU call of the round).

Inputs:
  gpurun_out/r2_final_launches_ncu.csv   launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras`
  gpurun_out/r2_final_prof_launches.csv  launch list of profiles/profile_run.py 1.0 (one monolithic pass per step)
  gpurun_out/r2_final_full_raw.csv       `ncu -i r2_final_full.ncu-rep --page raw --csv` of profiles/profile_run.py 1.0
  gpurun_out/r2_final_prof_plain.log     stdout of the un-profiled profiles/profile_run.py (unit counts)
  gpurun_out/r2_final_bench.json         the un-profiled `python bench.py --steps 20 --warmup 5`
Usage: python profiles/make_summary_r2.py
"""
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = lambda *p: os.path.join(ROOT, "gpurun_out", *p)
P = lambda *p: os.path.join(ROOT, "profiles", *p)


def short_name(s):
    s = re.sub(r"\(.*", "", s).replace("<unnamed>::", "").replace("void ", "").replace("rl::", "")
    return re.sub(r"<.*", "", s) or "(memset/memcpy)"


def launch_table(path, top=24):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        a = agg.setdefault(short_name(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    own = sum(a[1] for k, a in agg.items() if k.startswith("k_"))
    lines = ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        lines.append("| `%s` | %d | %.1f | %.1f%% |" % (k, c, t, 100 * t / tot))
    return len(rows) - 1, 100 * own / tot, lines, agg, tot


out = ["# Round 2 - ncu evidence, final state of the round\n",
       "Commands: `profiles/r2_final_commands.sh` (one gpurun call, 1 B200; every ncu pass follows the same command exiting 0 "
       "without ncu). Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n"]

for f, dst in (("r2_final_bench.json", "r2_final_bench.json"), ("r2_final_bench_reference.json", "r2_final_bench_reference.json"),
               ("r2_final_launches_ncu.csv", "r2_final_launches_ncu.csv"), ("r2_final_prof_launches.csv", "r2_launches_full.csv")):
    if os.path.exists(G(f)) and os.path.getsize(G(f)) > 0:
        shutil.copy(G(f), P(dst))
d = json.loads([l for l in open(P("r2_final_bench.json")) if l.startswith("{")][-1])

# ---------------------------------------------------------------- launch list of the bench command
n, own, lines, agg_b, tot_b = launch_table(P("r2_final_launches_ncu.csv"))
out.append("## Launch list of the bench command (profiles/r2_final_launches_ncu.csv)\n")
out.append("`ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline "
           "--no-extras` (C3 at full scale; the verification pass, the device-resident steps and the end-to-end steps through the eager "
           "chunk pipeline all appear). %d launches, %.1f %% of the kernel time in this repo's own kernels (the rest: cub radix sort / "
           "scan / run-length passes and memsets).\n" % (n, own))
out += lines
km = d["kernels_ms_per_step"]
out.append("\nCUDA-event stage times of the un-profiled run (profiles/r2_final_bench.json), ms per device-resident step of %.2f ms: "
           % d["ms_per_step"] + ", ".join("%s %.2f" % (k, v) for k, v in sorted(km.items(), key=lambda kv: -kv[1])) + ".")
rf = d["roofline"]
out.append("Roofline kernel `%s`: event share of the step %.1f %%; ncu share of this repo's kernels in the monolithic pass below: see table."
           % (rf.get("kernel", "k_pcr"), 100 * km.get("pcr", 0) / d["ms_per_step"]))

# ---------------------------------------------------------------- launch list of the monolithic pass
n, own, lines, agg_p, tot_p = launch_table(P("r2_launches_full.csv"))
out.append("\n## Launch list of one monolithic pass at full scale (profiles/r2_launches_full.csv)\n")
out.append("`ncu --metrics gpu__time_duration.sum --clock-control none --csv python profiles/profile_run.py 1.0` (eager pipeline off, one "
           "chunk: every stage is one launch at full size; two passes). %d launches, %.1f %% own kernels.\n" % (n, own))
out += lines
share = {"pcr": "k_pcr", "fragment": "k_fragment", "intervals": "k_intervals", "prefix": "k_prefix", "expand": "k_expand"}
ev_tot = sum(km.get(k, 0) for k in share)
nc_tot = sum(agg_p.get(v, [0, 0])[1] for v in share.values())
out.append("\nShare check (five largest own kernels; event share of their sum vs ncu share of their sum): " + ", ".join(
    "%s %.0f %% / %.0f %%" % (v, 100 * km.get(k, 0) / ev_tot, 100 * agg_p.get(v, [0, 0])[1] / nc_tot) for k, v in share.items()) + ".")

# ---------------------------------------------------------------- full capture
rows = list(csv.reader(open(G("r2_final_full_raw.csv"))))
hdr, units = rows[0], rows[1]
g = lambda r, n: r[hdr.index(n)]
u = lambda n: units[hdr.index(n)]
fl = lambda r, n: float(g(r, n).replace(",", "") or 0) if n in hdr else float("nan")
st = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
best = {}
for r in rows[2:]:
    name = short_name(g(r, "Kernel Name"))
    if name not in best or fl(best[name], "gpu__time_duration.sum") < fl(r, "gpu__time_duration.sum"):
        best[name] = r
out.append("\n## `ncu --set full --import-source on --clock-control none` of `python profiles/profile_run.py 1.0`, largest launch of each kernel\n")
out.append("| kernel | time %s | DRAM read %s | DRAM write %s | DRAM %% peak | L2 hit %% | issue active %% | FP64 pipe %% | regs | warps active %% | active lanes / inst | top stalls |"
           % (u("gpu__time_duration.sum"), u("dram__bytes_read.sum"), u("dram__bytes_write.sum")))
out.append("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for name, r in sorted(best.items(), key=lambda kv: -fl(kv[1], "gpu__time_duration.sum")):
    t = sum(fl(r, h) for h in st) or 1.0
    top = sorted(((fl(r, h), h) for h in st), reverse=True)[:3]
    tops = ", ".join("%s %.0f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / t) for v, h in top)
    out.append("| `%s` | %.1f | %.3f | %.3f | %.1f | %.1f | %.1f | %.1f | %s | %.1f | %.1f | %s |" % (
        name, fl(r, "gpu__time_duration.sum"), fl(r, "dram__bytes_read.sum"), fl(r, "dram__bytes_write.sum"),
        fl(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), fl(r, "lts__t_sector_hit_rate.pct"),
        fl(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        fl(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), g(r, "launch__registers_per_thread"),
        fl(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), fl(r, "smsp__thread_inst_executed_per_inst_executed.ratio"), tops))

# ---------------------------------------------------------------- DRAM traffic per unit (bench.py: roofline.traffic)
m = re.search(r"nt (\d+) molecules (\d+) fragments (\d+) species (\d+) sampled (\d+)", open(G("r2_final_prof_plain.log")).read())
nt, mol, frag, sp, smp = (int(x) for x in m.groups())
units_of = {"k_pcr": ("pcr", sp, "species"), "k_intervals": ("intervals", frag, "queued interval (~ kept fragment)"),
            "k_prefix": ("prefix", nt, "nt (expressed)"), "k_fragment": ("fragment", mol, "molecule"),
            "k_expand": ("expand", smp, "sampled fragment"), "k_ingest": ("ingest", nt, "nt")}
to_bytes = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full of profiles/profile_run.py "
                       "(C3 at full scale, whole-library launches; round 2 final), divided by the units of that launch; bench.py scales it to the run"}
for name, (key, n_units, unit) in units_of.items():
    if name in best:
        r = best[name]
        b = fl(r, "dram__bytes_read.sum") * to_bytes[u("dram__bytes_read.sum")] + fl(r, "dram__bytes_write.sum") * to_bytes[u("dram__bytes_write.sum")]
        traffic[key] = {"bytes_per_unit": round(b / n_units, 3), "unit": unit, "launch_bytes": int(b), "launch_units": n_units}
json.dump(traffic, open(P("ncu_traffic.json"), "w"), indent=1)
out.append("\nUnits of the pass: %d nt, %d molecules, %d fragments, %d species, %d sampled fragments." % (nt, mol, frag, sp, smp))
out.append("DRAM traffic per unit of work (profiles/ncu_traffic.json): " + ", ".join(
    "%s %.1f B/%s" % (k, v["bytes_per_unit"], v["unit"]) for k, v in traffic.items() if k != "_comment") + ".")
out.append("Achieved algorithmic GB/s per kernel group in the un-profiled step (bench.py, DESIGN.md section 5): " + ", ".join(
    "%s %.0f" % kv for kv in d.get("kernels_gbs", {}).items()) + ".")

# ---------------------------------------------------------------- headline numbers
e = d["e2e"]
out.append("\n## Headline numbers of the same call (profiles/r2_final_bench.json)\n")
out.append("Device-resident step %.2f ms = %.3g %s; end to end through the C-ABI with pinned host arrays %.2f ms (%.3g); roofline `%s`: "
           "%.1f %s of %.0f (%s), frac %.4f, ncu traffic %s." % (
               d["ms_per_step"], d["value"], d["unit"], e.get("ms_per_step", float("nan")), e["value"], rf.get("kernel", "k_pcr"),
               rf["achieved"], rf["unit"], rf["peak"], rf.get("peak_source", ""), rf["frac"], rf.get("traffic")))
out.append("- copy ceiling of this box for the same bytes (pinned H2D + D2H, nothing else): %.2f ms; e2e = %.2f of it" % (
    e.get("copy_ceiling_ms", float("nan")), e.get("copy_ceiling_ms", float("nan")) / e["ms_per_step"]))
for k in ("e2e_pageable", "e2e_packed_in", "e2e_text_in", "e2e_fasta", "e2e_dropin"):
    if isinstance(d.get(k), dict) and "ms_per_step" in d[k]:
        out.append("- `%s`: %.2f ms per step" % (k, d[k]["ms_per_step"]))
if os.path.exists(P("r2_final_bench_reference.json")):
    try:
        ref = json.loads([l for l in open(P("r2_final_bench_reference.json")) if l.startswith("{")][-1])
        out.append("- reference arm (profiles/r2_final_bench_reference.json): %.3g %s on %s cores, sample: %s" % (
            ref["value"], ref["unit"], ref["cpu_baseline"]["cores"], ref["cpu_baseline"]["sample"]))
    except Exception as ex:  # noqa: BLE001
        out.append("- reference arm line unreadable: %r" % (ex,))
open(P("r2_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
