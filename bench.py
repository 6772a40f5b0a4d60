#!/usr/bin/env python3
"""bench.py - fragments/sec through the library-construction hot path (frag + PCR + size-select).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scale S]

Workload (config C3 of BASELINE.json / SURVEY.md 8d): synthetic human-transcriptome-scale set - 200 000 transcripts
(lognormal lengths, ~4.5e8 nt), ~1e7 molecules (sel gamma-mixture expression), 1e7 requested fragments, 15 PCR cycles,
-d '0.9:sn:(600,50,4,300,2000) + 0.1:n:(700,1,600,2000)' -f after_prim_double -eg '(1.0,0.5,0.8)' -el '(0.1,0.7,1.0)'.
One STEP = one pass of the whole path over that library: target lengths -> priming prefix sums -> fragmentation ->
species dedup -> PCR -> per-length pools -> without-replacement draw -> fragment records.

    value : whole-job fragments/s with the transcripts already resident in HBM when the timed region starts
    e2e   : the same through the public API (GpuPool / GpuSampler over the C ABI) with HOST buffers: pinned-host ->
            device copy of the transcripts and device -> pinned-host read of the sampled fragment records inside
            the timed region, every step
    N > 1 : weak scaling - every rank owns its own C3-shaped shard (rank r: transcripts [r*200k, (r+1)*200k)),
            N x 1e7 fragments are requested from the global pool; the only collective is the all-gather of
            per-length pool totals.

--impl reference times the reference's own release binary (oracle/_ref) on the box's host cores on a bounded
sample of the same workload.  The oracle / reference binary are used here only as the CPU baseline legs.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3_FLAGS = ["-d", "0.9:sn:(600,50,4,300,2000) + 0.1:n:(700,1,600,2000)", "-f", "after_prim_double", "-c", "15",
            "-eg", "(1.0,0.5,0.8)", "-el", "(0.1,0.7,1.0)"]
C3_TRANSCRIPTS, C3_MOLECULES, C3_FRAGS = 200_000, 10_000_000, 10_000_000
REF_SAMPLE_TRANSCRIPTS = 4000  # bounded sample for the CPU reference: 1/50 of C3
SEED = 20261019


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip() == "Active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def make_args(n_frags, seed):
    from rlsim_b200.args import CmdArgs
    return CmdArgs.parse(["-n", str(n_frags), "-si", str(seed)] + C3_FLAGS + ["-"])


def reference_arm(opts, rank):
    """The reference's own CPU implementation (release binary) on a bounded sample of the workload."""
    from oracle import refbin
    from rlsim_b200 import synth
    if rank != 0:
        return
    n_tr = max(50, int(REF_SAMPLE_TRANSCRIPTS * opts.scale))
    frac = n_tr / C3_TRANSCRIPTS
    bases, off, levels = synth.transcriptome(n_tr, seed=SEED, total_molecules=C3_MOLECULES * frac)
    n_frags = int(C3_FRAGS * frac)
    cores = os.cpu_count() or 1
    have_bin = refbin.available()
    sample = ("first %d of the %d C3 transcripts' distribution (%d molecules), -n %d, same flags; %s -t %d"
              % (n_tr, C3_TRANSCRIPTS, int(levels.sum()), n_frags, "release binary rlsim 1.4" if have_bin else "oracle port (ref mode)", cores))
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "sample.fas")
        synth.write_fasta(fa, bases, off, levels)
        times, sampled = [], 0
        for step in range(opts.warmup + opts.steps):
            t0 = time.perf_counter()
            if have_bin:
                r = refbin.run(["-n", str(n_frags), "-si", str(step + 1)] + C3_FLAGS, fa, threads=cores, keep_fasta=False)
                sampled = sum(r["report"]["Sampled fragments"]["Data"].values())
            else:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import util
                from oracle import pyoracle
                a = make_args(n_frags, step + 1)
                seqs = [bases[int(off[i]):int(off[i + 1])].tobytes() for i in range(n_tr)]
                s = util.run_oracle(pyoracle, a, None, seqs, [int(x) for x in levels], pyoracle.MODE_REF)
                sampled = len(s.sampled()["tr"])
            dt = time.perf_counter() - t0
            if step >= opts.warmup:
                times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    value = sampled / (ms / 1000.0)
    line = {"impl": "reference", "metric": "fragments/sec (frag+PCR+size-select)", "value": value, "unit": "fragments/s",
            "n_gpus": opts.gpus, "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64+u64", "data": "synthetic",
            "config": {"workload": "C3 sample: " + sample},
            "cpu_baseline": {"value": value, "unit": "fragments/s", "cores": cores,
                             "kind": "reference" if have_bin else "port", "sample": sample},
            "e2e": {"value": value, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the C3 workload per GPU (1.0 = the named config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fasta", action="store_true", help="skip the extra end-to-end measurement that includes FASTA text")
    opts = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if opts.impl == "reference":
        reference_arm(opts, rank)
        return
    # libraries (NCCL) print banners on stdout: keep fd 1 clean for the single JSON line
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from rlsim_b200 import _native, shard, synth
    from rlsim_b200.pool import GpuPool, GpuSampler

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    cpus = shard.bind_to_gpu_numa_node(local_rank) if world > 1 else None  # before any pinned allocation
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic library of this rank (weak scaling: every rank owns a C3-shaped shard)
    n_tr = max(50, int(C3_TRANSCRIPTS * opts.scale))
    n_mol = C3_MOLECULES * opts.scale
    n_frags_global = int(C3_FRAGS * opts.scale) * world
    bases_np, off_np, levels_np = synth.transcriptome(n_tr, seed=SEED + rank, total_molecules=n_mol)
    bases = torch.from_numpy(bases_np).pin_memory()
    off = torch.from_numpy(off_np.view(np.int64)).pin_memory()
    levels = torch.from_numpy(levels_np.view(np.int64)).pin_memory()
    first = rank * n_tr
    h2d_bytes = bases.numel() + off.numel() * 8 + levels.numel() * 8

    pool = GpuPool(local_rank, rank, world, dev if world > 1 else None)
    sampler = GpuSampler()
    h = pool.handle
    out_cap = int(C3_FRAGS * opts.scale * 1.5) + 1024
    out = dict(tr=torch.empty(out_cap, dtype=torch.int32).pin_memory(), start=torch.empty(out_cap, dtype=torch.int32).pin_memory(),
               end=torch.empty(out_cap, dtype=torch.int32).pin_memory(), minus=torch.empty(out_cap, dtype=torch.uint8).pin_memory(),
               id=torch.empty(out_cap, dtype=torch.int64).pin_memory())

    exchange = shard.RecordExchange(dev) if world > 1 else None

    def select(step):
        if world > 1:  # classes partitioned over ranks + one all-to-all of the drawn copy ranks
            exchange.run(h, rank, world, n_frags_global // world + 4096, pool=pool)
        else:
            h.sample()

    e2e_phase = {"configure+target": 0.0, "h2d (add_transcripts)": 0.0, "build_library": 0.0, "select": 0.0, "d2h (records)": 0.0}

    def step_e2e(step, timed=False):
        """public API, host buffers in, host records out"""
        t = [time.perf_counter()]
        args = make_args(n_frags_global, step + 1)
        pool.reset().configure(args)
        t.append(time.perf_counter())
        h.set_first_index(first)
        h.add_transcripts_raw(bases, off, levels, n_tr)
        t.append(time.perf_counter())
        h.build_library()
        t.append(time.perf_counter())
        select(step)
        t.append(time.perf_counter())
        n = int(h.stats().n_sampled)
        if n > out_cap:
            raise SystemExit("output buffer too small")
        h.sampled(into=out)
        t.append(time.perf_counter())
        if timed:
            for k, a, b in zip(e2e_phase, t[:-1], t[1:]):
                e2e_phase[k] += 1000.0 * (b - a)
        return n

    def step_resident(step):
        """transcripts stay in HBM; everything derived from them is recomputed"""
        args = make_args(n_frags_global, step + 1)
        pool.configure(args)  # new seeds; invalidates prefix sums / species / pools
        h.build_library()
        select(step)
        return int(h.stats().n_sampled)

    # ---- warm-up (also allocates every device buffer)
    for s in range(opts.warmup):
        step_e2e(s)
    clocks = ClockSampler(local_rank)
    clocks.start()  # sampled across both timed regions
    time.sleep(0.3)
    # ---- e2e
    barrier()
    t0 = time.perf_counter()
    n_e2e = 0
    for s in range(opts.steps):
        n_e2e += step_e2e(100 + s, timed=True)
    barrier()
    t_e2e = time.perf_counter() - t0
    d2h_bytes = (n_e2e // max(opts.steps, 1)) * 21 + 7 * 8 * (int(h.stats().high) + 2)
    e2e_eager = bool(int(h.stats().flags) & _native.F_EAGER)
    # ---- e2e including FASTA text (reported separately; frag.go:71-126 formatted on the device)
    e2e_fasta = None
    if world == 1 and not opts.no_fasta:
        names = [b"T%d$%d" % (i, int(levels_np[i])) for i in range(n_tr)]
        names_u8 = np.frombuffer(b"".join(names), dtype=np.uint8).copy()
        names_off = np.zeros(n_tr + 1, dtype=np.uint64)
        names_off[1:] = np.cumsum([len(x) for x in names])
        text_bytes = h.fasta_into(names_u8, names_off, None)  # records of the last e2e step: sizes the pinned buffer
        text = torch.empty(int(text_bytes * 1.05) + (1 << 20), dtype=torch.uint8).pin_memory()
        fsteps = max(1, min(3, opts.steps))
        ms_f0 = h.timings()[0]["format"]
        for timed in (False, True):
            barrier()
            t0 = time.perf_counter()
            nf = tb = 0
            for s in range(fsteps):
                nf += step_e2e(300 + s)
                tb += h.fasta_into(names_u8, names_off, text)
            barrier()
            t_f = time.perf_counter() - t0
        e2e_fasta = {"value": nf / t_f, "unit": "fragments/s", "steps": fsteps, "ms_per_step": 1000.0 * t_f / fsteps,
                     "text_bytes_per_step": tb // fsteps, "format_kernels_ms_per_step": (h.timings()[0]["format"] - ms_f0) / (2 * fsteps),
                     "note": "the e2e step above plus rlsim_format_fasta: records formatted on the device, text copied to pinned host memory"}
        del text
    # ---- e2e from raw FASTA text: records split / despaced / upper-cased / parsed on the device
    e2e_text_in = None
    if world == 1 and not opts.no_fasta:
        raw = torch.from_numpy(synth.fasta_text(bases_np, off_np, levels_np, width=70).copy()).pin_memory()
        tsteps = max(1, min(5, opts.steps))

        def step_text(step):
            pool.reset().configure(make_args(n_frags_global, step + 1))
            h.set_first_index(0)
            idx = h.add_fasta(raw, 1.0)
            assert idx["n_added"] == n_tr
            h.build_library()
            select(step)
            n = int(h.stats().n_sampled)
            h.sampled(into=out)
            return n

        step_text(400)
        barrier()
        t0 = time.perf_counter()
        nt_in = 0
        for s in range(tsteps):
            nt_in += step_text(401 + s)
        barrier()
        t_t = time.perf_counter() - t0
        e2e_text_in = {"value": nt_in / t_t, "unit": "fragments/s", "steps": tsteps, "ms_per_step": 1000.0 * t_t / tsteps,
                       "text_bytes_per_step": int(raw.numel()),
                       "note": "the e2e step fed with the raw FASTA text (70-column lines) through rlsim_add_fasta instead of parsed arrays"}
        del raw
    # ---- device-resident
    for s in range(min(opts.warmup, 2)):
        step_resident(50 + s)
    ms0, nl0 = h.timings()
    barrier()
    t0 = time.perf_counter()
    n_res = 0
    for s in range(opts.steps):
        n_res += step_resident(200 + s)
    barrier()
    t_res = time.perf_counter() - t0
    clk = clocks.stop()
    ms1, nl1 = h.timings()
    st = h.stats()

    if world > 1:
        t = torch.tensor([t_res, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(t[0]), float(t[1])
        c = torch.tensor([n_res, n_e2e], dtype=torch.int64, device=dev)
        dist.all_reduce(c)
        n_res, n_e2e = int(c[0]), int(c[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    K = max(opts.steps, 1)
    stage_ms = {k: (ms1[k] - ms0[k]) / K for k in ms1}
    launches = sum(nl1[k] - nl0[k] for k in nl1) // K
    mirrored = bool(int(st.flags) & _native.F_MIRRORED_REVERSE)
    expressed = levels_np > 0  # pool.go:104: unexpressed transcripts get no prefix sums
    lens = (off_np[1:] - off_np[:-1]).astype(np.int64)
    nt_total = int(lens[expressed].sum()) + int(expressed.sum()) * int(st.polya_max)
    # algorithmic bytes per step of each own kernel group
    n_intervals = float(st.n_fragments) * 1.0  # queued intervals ~ kept fragments at this -d (upper bound of what the searches touch)
    alg = {
        # 0.25 B/nt read + per strand (4 B fine + 8/64 B coarse) written, expressed transcripts; one strand when the reverse
        # profile is read mirrored from the forward sums (symmetric k-mer table)
        "prefix": (0.25 + (1 if mirrored else 2) * 4.125) * nt_total,
        "fragment": 8.0 * float(st.n_molecules) + 24.0 * n_intervals,   # k_fragment: 8 B/molecule in, 24 B per queued interval out
        "intervals": (24.0 + 2 * 12 * 8.0 + 8.0) * n_intervals,         # k_intervals: record in, 2 searches x ~12 probes x 8 B, key out
        "pcr": 32.0 * float(st.n_species),                # 32 B/species
        "expand": 150.0 * float(st.n_sampled),            # ~150 B/sampled fragment
    }
    peak, peak_src = peaks()
    kernels = {}
    for k, b in alg.items():
        if stage_ms.get(k, 0) > 0:
            kernels[k] = {"ms": stage_ms[k], "alg_bytes": b, "gbs": b / (stage_ms[k] * 1e6)}
    for k in ("dedup", "select", "target", "layout"):
        kernels[k] = {"ms": stage_ms[k]}
    own = {k: v for k, v in kernels.items() if "gbs" in v}
    dom = max(own, key=lambda k: own[k]["ms"])
    # measured DRAM traffic per unit from the committed ncu --set full capture (profiles/ncu_traffic.json), scaled to this run
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        units = {"pcr": float(st.n_species), "prefix": float(nt_total), "intervals": n_intervals, "fragment": float(st.n_molecules),
                 "expand": float(st.n_sampled)}
        if dom in tj and dom in units:
            traffic = tj[dom]["bytes_per_unit"] * units[dom]
    roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": own[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": own[dom]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                "note": ("k_pcr is FP64-issue bound (15 exact binomial draws per 32 B species): see profiles/r1_final_ncu_summary.md"
                         if dom == "pcr" else ""),
                "all": {("k_" + k): {"ms": round(v["ms"], 4), "achieved": round(v["gbs"], 2), "frac": round(v["gbs"] / peak, 4)} for k, v in own.items()},
                "all_note": "k_prefix shares the GPU with k_target (own stream) in this step: alone it takes 0.70 ms = 2.7 TB/s "
                            "(profiles/r1_final_ncu_summary.md)"}
    ms_per_step = 1000.0 * t_res / K
    line = {
        "metric": "fragments/sec (frag+PCR+size-select)", "value": n_res / t_res, "unit": "fragments/s", "n_gpus": world,
        "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64+u64", "data": "synthetic",
        "config": {"workload": "C3 synthetic human-transcriptome-scale: %d transcripts / %d nt / %d molecules per GPU, %d requested "
                               "fragments in total, 15 cycles, after_prim_double" % (n_tr, int(bases.numel()), int(st.n_molecules), n_frags_global),
                   "flags": " ".join(C3_FLAGS), "scale": opts.scale, "sharding": "contiguous transcripts per rank; allgather of per-length totals + all-to-all of drawn copy ranks",
                   "cpu_binding_rank0": (sorted(cpus)[0], sorted(cpus)[-1], len(cpus)) if cpus else None,
                   "l2": "inputs (0.45 GB bases + %.1f GB prefix sums per GPU) exceed the 126 MB L2; no explicit flush" % ((1 if mirrored else 2) * 4.125 * nt_total / 1e9),
                   "reverse_profile": "mirrored forward sums" if mirrored else "explicit arrays",
                   "e2e_path": "eager chunk pipeline (fragment+dedup+PCR per arrived group of 32 MB chunks)" if e2e_eager else "monolithic",
                   "species_per_gpu": int(st.n_species), "fragments_after_fragmentation_per_gpu": int(st.n_fragments),
                   "pool_total_rank0": int(st.pool_total)},
        "e2e": {"value": n_e2e / t_e2e, "unit": "fragments/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": 1000.0 * t_e2e / K, "phases_ms_rank0": {k: round(v / K, 3) for k, v in e2e_phase.items()}},
        "e2e_fasta": e2e_fasta,
        "e2e_text_in": e2e_text_in,
        "stage_rates": {  # per-stage rates of the device-resident step
            "prefix_nt_per_s": nt_total / (stage_ms["prefix"] * 1e-3) if stage_ms.get("prefix", 0) > 0 else None,
            "fragmentation_molecules_per_s": float(st.n_molecules) / ((stage_ms["fragment"] + stage_ms.get("intervals", 0.0)) * 1e-3),
            "pcr_species_per_s": float(st.n_species) / (stage_ms["pcr"] * 1e-3),
            "selection_fragments_per_s": float(st.n_sampled) / ((stage_ms["select"] + stage_ms["expand"]) * 1e-3),
        },
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "kernels_ms_per_step": {k: round(v["ms"], 4) for k, v in kernels.items()},
        "kernels_gbs": {k: round(v["gbs"], 2) for k, v in own.items()},
    }
    if world == 1 and not opts.no_cpu_baseline:
        # the reference's release binary on this box's host cores, bounded sample (reported baseline, not the target)
        try:
            import io
            from contextlib import redirect_stdout
            buf = io.StringIO()
            ref_opts = argparse.Namespace(gpus=1, steps=1, warmup=0, scale=opts.scale)
            with redirect_stdout(buf):
                reference_arm(ref_opts, 0)
            line["cpu_baseline"] = json.loads(buf.getvalue())["cpu_baseline"]
        except Exception as e:  # the baseline must never break the GPU line
            line["cpu_baseline"] = {"value": None, "unit": "fragments/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": "failed: %s" % e}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
