#!/usr/bin/env python3
"""bench.py - fragments/sec through the library-construction hot path (frag + PCR + size-select).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scale S] [--scaling weak|strong]

Workload (config C3 of BASELINE.json / SURVEY.md 8d): synthetic human-transcriptome-scale set - 200 000 transcripts
(lognormal lengths, ~4.5e8 nt), ~1e7 molecules (sel gamma-mixture expression), 1e7 requested fragments, 15 PCR cycles,
-d '0.9:sn:(600,50,4,300,2000) + 0.1:n:(700,1,600,2000)' -f after_prim_double -eg '(1.0,0.5,0.8)' -el '(0.1,0.7,1.0)'.
One STEP = one pass of the whole path over that library: target lengths -> priming prefix sums -> fragmentation ->
species dedup -> PCR -> per-length pools -> without-replacement draw -> fragment records.

    value        : whole-job fragments/s with the transcripts already resident in HBM when the timed region starts
    e2e          : the same through the public API (GpuPool / GpuSampler over the C ABI) with HOST buffers: pinned-host ->
                   device copy of the transcripts and device -> pinned-host read of the sampled fragment records (packed
                   form, 6 B per record) inside the timed region, every step
    e2e_pageable : e2e with plain (unpinned) numpy buffers on both sides - what a cgo caller hands over
    e2e_packed_in: e2e from a host that keeps the transcriptome as 2-bit codes + N mask (rlsim_add_transcripts_packed)
    e2e_dropin   : FASTA text in -> FASTA text out (rlsim_add_fasta ... rlsim_format_fasta): what the reference's CLI does
    verified     : outside the timed regions - the eager end-to-end result, the pageable one and the monolithic
                   device-resident one give the same digest for one seed
    N > 1        : weak scaling (every rank owns its own C3-shaped shard, N x 1e7 fragments requested from the global
                   pool) is the headline; the line also carries `strong`: the north_star configuration itself - ONE C3,
                   10 M fragments, transcripts sharded N ways.  The exchange (all-gather of per-length totals, routing of
                   the drawn copy ranks) runs inside the library over NCCL on the handle's stream.

--impl reference times the reference's own release binary (oracle/_ref) on the box's host cores on a bounded sample of
the same workload (1/10-scale C3 by default: 20 000 transcripts, 1e6 molecules, -n 1e6, as SURVEY 8d prescribes when the
full run - ~10-17 min per step - does not fit).  The oracle / reference binary are used here only as the CPU baseline legs.
"""
import argparse
import hashlib
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
REF_SCALE = 0.1          # bounded sample for the CPU reference: 1/10 of C3
REF_TIME_BUDGET_S = 150  # the reference arm stops starting new steps after this many seconds
SEED = 20261019
METRIC = "fragments/sec (frag+PCR+size-select)"


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


def reference_sample(scale):
    """The bounded sample both arms can run: the first scale*200000 transcripts' worth of the C3 distribution."""
    from rlsim_b200 import synth
    n_tr = max(50, int(C3_TRANSCRIPTS * scale))
    frac = n_tr / C3_TRANSCRIPTS
    bases, off, levels = synth.transcriptome(n_tr, seed=SEED, total_molecules=C3_MOLECULES * frac)
    return n_tr, int(C3_FRAGS * frac), bases, off, levels


def run_reference(scale, steps, warmup, budget_s):
    """-> dict with the reference's fragments/s on the sample, its per-stage rates (from the -v log timestamps:
    fragstats.go:132-135, sampler.go:152-153) and what was run."""
    from oracle import refbin
    from rlsim_b200 import synth
    n_tr, n_frags, bases, off, levels = reference_sample(scale)
    cores = os.cpu_count() or 1
    have_bin = refbin.available()
    sample = ("%.3g-scale C3: first %d of the %d transcripts' distribution (%d molecules), -n %d, same flags; %s -t %d; FASTA file in, "
              "FASTA text out to /dev/null" % (n_tr / C3_TRANSCRIPTS, n_tr, C3_TRANSCRIPTS, int(levels.sum()), n_frags,
                                             "release binary rlsim 1.4" if have_bin else "oracle port (ref mode)", cores))
    times, stage, sampled, species = [], [], 0, None
    t_begin = time.perf_counter()
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "sample.fas")
        synth.write_fasta(fa, bases, off, levels)
        done = 0
        warmup = min(warmup, 1)  # a CPU binary: one untimed run warms the page cache, more only burn the time budget
        for step in range(warmup + steps):
            if done > warmup and time.perf_counter() - t_begin > budget_s:
                break
            t0 = time.perf_counter()
            if have_bin:
                r = refbin.run(["-n", str(n_frags), "-si", str(step + 1)] + C3_FLAGS, fa, threads=cores, keep_fasta=False)
                sampled = sum(r["report"]["Sampled fragments"]["Data"].values())
                ts = refbin.stage_times(r["log"])
                if {"frag_start", "frag_done", "sample_done"} <= set(ts):
                    stage.append((ts["frag_done"] - ts["frag_start"], ts["sample_done"] - ts["frag_done"]))
            else:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import util
                from oracle import pyoracle
                a = make_args(n_frags, step + 1)
                seqs = [bases[int(off[i]):int(off[i + 1])].tobytes() for i in range(n_tr)]
                s = util.run_oracle(pyoracle, a, None, seqs, [int(x) for x in levels], pyoracle.MODE_REF)
                sampled = len(s.sampled()["tr"])
            dt = time.perf_counter() - t0
            if step >= warmup:
                times.append(dt)
            done += 1
    ms = 1000.0 * float(np.mean(times))
    out = {"value": sampled / (ms / 1000.0), "ms_per_step": ms, "steps_done": len(times), "warmup_done": done - len(times),
           "cores": cores, "kind": "reference" if have_bin else "port", "sample": sample, "scale": n_tr / C3_TRANSCRIPTS,
           "n_transcripts": n_tr, "n_molecules": int(levels.sum()), "n_sampled": int(sampled)}
    if stage:
        fr, sm = float(np.mean([s[0] for s in stage])), float(np.mean([s[1] for s in stage]))
        out["stages"] = {"frag_pcr_s": fr, "sampling_print_s": sm,
                         "fragmentation_molecules_per_s": int(levels.sum()) / fr if fr > 0 else None,
                         "selection_fragments_per_s": sampled / sm if sm > 0 else None,
                         "note": "from the -v log timestamps: 'Fragmenting transcripts' -> 'Initialized' -> 'Total number of fragments in the pool'"}
    return out


def reference_arm(opts, rank):
    """The reference's own CPU implementation (release binary) on a bounded sample of the workload."""
    if rank != 0:
        return
    scale = opts.scale if opts.scale_given else REF_SCALE
    r = run_reference(scale, opts.steps, opts.warmup, REF_TIME_BUDGET_S)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "fragments/s",
            "n_gpus": opts.gpus, "steps": r["steps_done"], "warmup": r["warmup_done"], "steps_requested": opts.steps,
            "warmup_requested": opts.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64+u64", "data": "synthetic",
            "config": {"workload": "C3 sample: " + r["sample"], "scale": r["scale"],
                       "bounded": "steps stop after %d s of wall time (a full-scale C3 step takes the reference 10-17 min: per-draw cost "
                                  "grows with the pool, random.go:439-475)" % REF_TIME_BUDGET_S},
            "cpu_baseline": {"value": r["value"], "unit": "fragments/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "stage_rates": r.get("stages"),
            "e2e": {"value": r["value"], "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scale", type=float, default=None, help="fraction of the C3 workload per GPU (1.0 = the named config)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1: headline of the line (the other one is reported beside it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the pageable / FASTA-text / strong-scaling extras")
    opts = ap.parse_args()
    opts.scale_given = opts.scale is not None
    if opts.scale is None:
        opts.scale = 1.0
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

    def pin(a):
        return torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).pin_memory()

    pool = GpuPool(local_rank, rank, world, dev if world > 1 else None)
    sampler = GpuSampler()
    h = pool.handle
    exchange = shard.RecordExchange(dev) if world > 1 and not pool.lib_comm else None

    class Workload:
        """One rank's share of a library + the host buffers of the end-to-end step."""

        def __init__(self, bases_np, off_np, levels_np, first, n_frags_global, label):
            self.bases_np, self.off_np, self.levels_np = bases_np, off_np, levels_np
            self.n_tr, self.first, self.n_frags_global, self.label = len(levels_np), first, n_frags_global, label
            self.bases, self.off, self.levels = pin(bases_np), pin(off_np), pin(levels_np)
            self.h2d_bytes = self.bases.numel() + self.off.numel() * 8 + self.levels.numel() * 8
            cap = int(n_frags_global / world * 1.6) + 4096
            self.cap = cap
            self.out = dict(tr_counts=torch.empty(self.n_tr, dtype=torch.int32).pin_memory(), start=torch.empty(cap, dtype=torch.int32).pin_memory(),
                            len_strand=torch.empty(cap, dtype=torch.int16).pin_memory())
            self.phase = {"configure+target": 0.0, "h2d (add_transcripts)": 0.0, "build_library": 0.0, "select": 0.0, "d2h (records)": 0.0}

    def select():
        if world > 1 and pool.lib_comm:
            h.sample_distributed()
        elif world > 1:  # torch.distributed fallback: classes partitioned over ranks + all-to-all of the drawn copy ranks
            exchange.run(h, rank, world, wl.n_frags_global // world + 4096, pool=pool)
        else:
            h.sample()

    def step_e2e(step, timed=False, pageable=None, packed=None):
        """public API, host buffers in, host records out.  pageable = (bases, off, levels, out) plain numpy buffers;
        packed = (codes, nmask) pinned arrays of rlsim_add_transcripts_packed."""
        t = [time.perf_counter()]
        pool.reset().configure(make_args(wl.n_frags_global, step + 1))
        t.append(time.perf_counter())
        h.set_first_index(wl.first)
        if pageable:
            h.add_transcripts_raw(pageable[0], pageable[1], pageable[2], wl.n_tr)
        elif packed:
            h.add_transcripts_packed(packed[0], packed[1], wl.off, wl.levels, wl.n_tr)
        else:
            h.add_transcripts_raw(wl.bases, wl.off, wl.levels, wl.n_tr)
        t.append(time.perf_counter())
        h.build_library()
        t.append(time.perf_counter())
        select()
        t.append(time.perf_counter())
        n = int(h.stats().n_sampled)
        if n > wl.cap:
            raise SystemExit("output buffer too small")
        h.sampled_packed(into=pageable[3] if pageable else wl.out, elem_bytes=2)
        t.append(time.perf_counter())
        if timed:
            for k, a, b in zip(wl.phase, t[:-1], t[1:]):
                wl.phase[k] += 1000.0 * (b - a)
        return n

    def step_resident(step):
        """transcripts stay in HBM; everything derived from them is recomputed"""
        pool.configure(make_args(wl.n_frags_global, step + 1))  # new seeds; invalidates prefix sums / species / pools
        h.build_library()
        select()
        return int(h.stats().n_sampled)

    def result_digest(packed, n):
        """sha256 over the species table (coordinates, initial and amplified copy numbers) and the sampled records"""
        sp = h.species()
        hs = hashlib.sha256()
        for k in ("tr", "start", "end", "count0", "count"):
            hs.update(np.ascontiguousarray(sp[k]).tobytes())
        tc = packed["tr_counts"]
        hs.update(np.ascontiguousarray(tc.numpy() if hasattr(tc, "numpy") else tc).tobytes())
        for k in ("start", "len_strand"):
            a = packed[k]
            hs.update(np.ascontiguousarray((a.numpy() if hasattr(a, "numpy") else a)[:n]).tobytes())
        return hs.hexdigest()

    def timed_loop(fn, steps, first_step, **kw):
        barrier()
        t0 = time.perf_counter()
        n = 0
        for s in range(steps):
            n += fn(first_step + s, **kw)
        barrier()
        return n, time.perf_counter() - t0

    def reduce_max_sum(t, n):
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            nn = torch.tensor([n], dtype=torch.int64, device=dev)
            dist.all_reduce(nn)
            return float(tt[0]), int(nn[0])
        return t, n

    # ---- synthetic library of this rank (weak scaling: every rank owns a C3-shaped shard)
    n_tr = max(50, int(C3_TRANSCRIPTS * opts.scale))
    n_frags_rank = int(C3_FRAGS * opts.scale)
    b_np, o_np, l_np = synth.transcriptome(n_tr, seed=SEED + rank, total_molecules=C3_MOLECULES * opts.scale)
    wl = weak = Workload(b_np, o_np, l_np, rank * n_tr, n_frags_rank * world, "weak")

    # ---- warm-up (also allocates every device buffer)
    for s in range(opts.warmup):
        step_e2e(s)
    clocks = ClockSampler(local_rank)
    clocks.start()  # sampled across the timed regions
    time.sleep(0.3)
    # ---- e2e
    n_e2e, t_e2e = timed_loop(step_e2e, opts.steps, 100, timed=True)
    n_last = int(h.stats().n_sampled)
    d2h_bytes = n_last * 6 + wl.n_tr * 4 + 7 * 8 * (int(h.stats().high) + 2)
    e2e_eager = bool(int(h.stats().flags) & _native.F_EAGER)
    e2e_groups = (int(h.stats().flags) >> _native.F_GROUPS_SHIFT) & 0xff

    # ---- verification (outside the timed regions): eager e2e == pageable e2e == monolithic resident, one seed
    verified, verify_note = None, None
    e2e_pageable = None
    if world == 1:
        VS = 7000
        n_v = step_e2e(VS)
        d_eager = result_digest(wl.out, n_v)
        eager_flag = bool(int(h.stats().flags) & _native.F_EAGER)
        os.environ["RLSIM_NO_GROUPS"] = "1"  # one monolithic launch per stage: the path that shares least with the eager one
        n_m = step_resident(VS)
        del os.environ["RLSIM_NO_GROUPS"]
        mono_flag = bool(int(h.stats().flags) & _native.F_EAGER)
        packed_m, _ = h.sampled_packed(elem_bytes=2)
        d_mono = result_digest(packed_m, n_m)
        verified = (d_eager == d_mono) and n_v == n_m
        verify_note = {"eager_e2e_digest": d_eager[:16], "monolithic_resident_digest": d_mono[:16], "n_sampled": n_v,
                       "eager_path": eager_flag, "resident_path_eager": mono_flag}
        if not opts.no_extras:
            # ---- e2e with pageable (unpinned) caller buffers: the library stages them through its pinned ring
            pg_out = dict(tr_counts=np.zeros(wl.n_tr, np.uint32), start=np.zeros(wl.cap, np.uint32), len_strand=np.zeros(wl.cap, np.uint16))
            pg = (wl.bases_np, wl.off_np, wl.levels_np, pg_out)
            n_p = step_e2e(VS, pageable=pg)
            staged = bool(int(h.stats().flags) & _native.F_PAGEABLE_STAGED)
            d_page = result_digest(pg_out, n_p)
            verified = verified and d_page == d_eager
            verify_note["pageable_e2e_digest"] = d_page[:16]
            psteps = max(1, min(10, opts.steps))
            step_e2e(VS + 1, pageable=pg)
            n_pg, t_pg = timed_loop(step_e2e, psteps, 500, pageable=pg)
            e2e_pageable = {"value": n_pg / t_pg, "unit": "fragments/s", "steps": psteps, "ms_per_step": 1000.0 * t_pg / psteps,
                            "staged_by_library": staged, "stage_threads": int(os.environ.get("RLSIM_STAGE_THREADS", "0")) or min(6, max(1, (os.cpu_count() or 2) - 1)),
                            "note": "same step with plain numpy buffers on both sides (what a cgo caller hands over): chunks staged through "
                                    "a ring of pinned slots by the library's copy threads"}
    # ---- e2e from a host that keeps its transcriptome packed (2-bit codes + N mask; 0.375 B per base on the host link)
    e2e_packed_in = None
    if not opts.no_extras:
        codes_np, nmask_np = _native.pack_bases(wl.bases_np)  # once, outside the timed region: the host's own resident format
        pk = (pin(codes_np), pin(nmask_np) if nmask_np is not None else None)
        if world == 1:
            n_k = step_e2e(VS, packed=pk)
            d_pk = result_digest(wl.out, n_k)
            verified = verified and d_pk == d_eager
            verify_note["packed_in_e2e_digest"] = d_pk[:16]
        ksteps = max(1, min(10, opts.steps))
        for s_ in range(2):
            step_e2e(7002 + s_, packed=pk)
        n_kk, t_kk = timed_loop(step_e2e, ksteps, 600, packed=pk)
        t_kk, n_kk = reduce_max_sum(t_kk, n_kk)  # whole job: max over ranks of the time, sum of the fragments
        e2e_packed_in = {"value": n_kk / t_kk, "unit": "fragments/s", "steps": ksteps, "ms_per_step": 1000.0 * t_kk / ksteps,
                         "h2d_bytes_per_step": int(pk[0].numel() + (pk[1].numel() if pk[1] is not None else 0) + wl.off.numel() * 8 + wl.levels.numel() * 8),
                         "note": "same step through rlsim_add_transcripts_packed: the host holds 2-bit codes (+ a mask for letters outside "
                                 "ACGT) instead of ASCII; packing is the host's one-off cost and is not timed"}
        del pk
    # ---- FASTA text in / out (SURVEY 8f rows 1 + 3): the reference CLI's own boundary
    e2e_text_in = e2e_dropin = None
    if world == 1 and not opts.no_extras:
        names = [b"T%07d" % i for i in range(wl.n_tr)]
        names_u8 = np.frombuffer(b"".join(names), dtype=np.uint8).copy()
        names_off = np.arange(wl.n_tr + 1, dtype=np.uint64) * 8
        raw = torch.from_numpy(synth.fasta_text(wl.bases_np, wl.off_np, wl.levels_np, width=70).copy()).pin_memory()
        text_buf = [None]

        def step_text(step, text_out=False):
            pool.reset().configure(make_args(wl.n_frags_global, step + 1))
            h.set_first_index(0)
            idx = h.add_fasta(raw, 1.0)
            assert idx["n_added"] == wl.n_tr
            h.build_library()
            select()
            n = int(h.stats().n_sampled)
            if text_out:
                if text_buf[0] is None:
                    size = h.fasta_into(names_u8, names_off, None)
                    text_buf[0] = torch.empty(int(size * 1.05) + (1 << 20), dtype=torch.uint8).pin_memory()
                step_text.bytes = h.fasta_into(names_u8, names_off, text_buf[0])
            else:
                h.sampled_packed(into=wl.out, elem_bytes=2)
            return n

        tsteps = max(1, min(5, opts.steps))
        step_text(400)
        n_t, t_t = timed_loop(step_text, tsteps, 401)
        e2e_text_in = {"value": n_t / t_t, "unit": "fragments/s", "steps": tsteps, "ms_per_step": 1000.0 * t_t / tsteps,
                       "text_bytes_per_step": int(raw.numel()),
                       "note": "the e2e step fed with the raw FASTA text (70-column lines) through rlsim_add_fasta instead of parsed arrays"}
        fsteps = max(1, min(3, opts.steps))
        step_text(450, text_out=True)
        n_d, t_d = timed_loop(step_text, fsteps, 451, text_out=True)
        e2e_dropin = {"value": n_d / t_d, "unit": "fragments/s", "steps": fsteps, "ms_per_step": 1000.0 * t_d / fsteps,
                      "text_in_bytes_per_step": int(raw.numel()), "text_out_bytes_per_step": int(step_text.bytes),
                      "note": "FASTA text in -> FASTA text out (rlsim_add_fasta ... rlsim_format_fasta, pinned buffers): the boundary of the "
                              "reference's CLI, which is what --impl reference times"}
        del raw
        text_buf[0] = None
    # ---- device-resident (the headline `value`)
    for s in range(min(opts.warmup, 2)):
        step_resident(50 + s)
    ms0, nl0 = h.timings()
    n_res, t_res = timed_loop(step_resident, opts.steps, 200)
    ms1, nl1 = h.timings()
    st = h.stats()
    weak_stats = dict(n_species=int(st.n_species), n_fragments=int(st.n_fragments), n_molecules=int(st.n_molecules),
                      pool_total=int(st.pool_total), n_sampled=int(st.n_sampled), polya_max=int(st.polya_max), flags=int(st.flags))

    # ---- copy ceiling: what this host can move for one e2e step (all ranks at once): pinned H2D + D2H of the same sizes
    ceiling_ms = None
    if not opts.no_extras:
        dbuf = torch.empty(wl.bases.numel(), dtype=torch.uint8, device=dev)
        dout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=dev)
        hout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8).pin_memory()

        def step_copy(step):
            dbuf.copy_(wl.bases, non_blocking=True)
            hout.copy_(dout, non_blocking=True)
            torch.cuda.synchronize()
            return 0

        step_copy(0)
        _, t_c = timed_loop(step_copy, 5, 0)
        ceiling_ms, _ = reduce_max_sum(1000.0 * t_c / 5, 0)
        del dbuf, dout, hout

    # ---- strong scaling (N > 1): the north_star configuration - ONE C3, 10 M fragments, transcripts sharded N ways
    strong = None
    if world > 1 and not opts.no_extras:
        fb, fo, fl = synth.transcriptome(n_tr, seed=SEED, total_molecules=C3_MOLECULES * opts.scale)  # the same set on every rank
        lens = (fo[1:] - fo[:-1]).astype(np.int64)
        bnd = shard.contiguous_partition(shard.work_estimate(lens, fl), world)
        a0, a1 = bnd[rank], bnd[rank + 1]
        sb = np.ascontiguousarray(fb[int(fo[a0]):int(fo[a1])])
        so = (fo[a0:a1 + 1] - fo[a0]).astype(np.uint64)
        del weak.bases, weak.off, weak.levels, weak.out
        wl = Workload(sb, so, fl[a0:a1].copy(), a0, n_frags_rank, "strong")
        del fb
        ssteps = max(1, min(10, opts.steps))
        for s in range(2):
            step_e2e(600 + s)
        ns_e, ts_e = timed_loop(step_e2e, ssteps, 610, timed=True)
        for s in range(2):
            step_resident(650 + s)
        ns_r, ts_r = timed_loop(step_resident, ssteps, 660)
        ts_e, ns_e = reduce_max_sum(ts_e, ns_e)
        ts_r, ns_r = reduce_max_sum(ts_r, ns_r)
        strong = {"scaling": "strong", "value": ns_r / ts_r, "unit": "fragments/s", "ms_per_step": 1000.0 * ts_r / ssteps, "steps": ssteps,
                  "e2e": {"value": ns_e / ts_e, "ms_per_step": 1000.0 * ts_e / ssteps, "h2d_bytes_per_step_rank0": int(wl.h2d_bytes),
                          "phases_ms_rank0": {k: round(v / ssteps, 3) for k, v in wl.phase.items()}},
                  "config": "C3 itself (BASELINE.json configs[2]): %d transcripts, %d requested fragments, contiguous shards balanced by "
                            "expression x length; rank 0 owns transcripts [%d, %d)" % (n_tr, n_frags_rank, a0, a1)}
    clk = clocks.stop()

    t_res, n_res = reduce_max_sum(t_res, n_res)
    t_e2e, n_e2e = reduce_max_sum(t_e2e, n_e2e)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    K = max(opts.steps, 1)
    stage_ms = {k: (ms1[k] - ms0[k]) / K for k in ms1}
    launches = sum(nl1[k] - nl0[k] for k in nl1) // K
    mirrored = bool(weak_stats["flags"] & _native.F_MIRRORED_REVERSE)
    expressed = weak.levels_np > 0  # pool.go:104: unexpressed transcripts get no prefix sums
    lens = (weak.off_np[1:] - weak.off_np[:-1]).astype(np.int64)
    nt_total = int(lens[expressed].sum()) + int(expressed.sum()) * weak_stats["polya_max"]
    # algorithmic bytes per step of each own kernel group
    n_intervals = float(weak_stats["n_fragments"])  # queued intervals ~ kept fragments at this -d
    alg = {
        # 0.25 B/nt read + per strand (4 B fine + 8/64 B coarse) written, expressed transcripts; one strand when the reverse
        # profile is read mirrored from the forward sums (symmetric k-mer table)
        "prefix": (0.25 + (1 if mirrored else 2) * 4.125) * nt_total,
        "fragment": 8.0 * float(weak_stats["n_molecules"]) + 24.0 * n_intervals,  # 8 B/molecule in, 24 B per queued interval out
        # record in (24 B) + key out (8 B) + two searches, each: P(start), P(end) and the hit entry = 3 x (8 B coarse + 4 B fine)
        "intervals": (24.0 + 8.0 + 2 * 3 * 12.0) * n_intervals,
        "pcr": 32.0 * float(weak_stats["n_species"]),       # 32 B/species
        # 21 B record + 6 B packed record out, two 8-byte reads of the species search that miss the cache
        "expand": (21.0 + 6.0 + 16.0) * float(weak_stats["n_sampled"]),
    }
    peak, peak_src = peaks()
    kernels = {}
    for k, b in alg.items():
        if stage_ms.get(k, 0) > 0:
            kernels[k] = {"ms": stage_ms[k], "alg_bytes": b, "gbs": b / (stage_ms[k] * 1e6)}
    for k in ("dedup", "select", "target", "layout", "exchange"):
        kernels[k] = {"ms": stage_ms.get(k, 0.0)}
    own = {k: v for k, v in kernels.items() if "gbs" in v}
    dom = max(own, key=lambda k: own[k]["ms"])
    # measured DRAM traffic per unit from the committed ncu --set full capture (profiles/ncu_traffic.json), scaled to this run
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        units = {"pcr": float(weak_stats["n_species"]), "prefix": float(nt_total), "intervals": n_intervals,
                 "fragment": float(weak_stats["n_molecules"]), "expand": float(weak_stats["n_sampled"])}
        if dom in tj and dom in units:
            traffic = tj[dom]["bytes_per_unit"] * units[dom]
    roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": own[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": own[dom]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                "note": ("k_pcr is FP64-issue bound (15 exact binomial draws per 32 B species): see profiles/ (ncu summary of this round)"
                         if dom == "pcr" else ""),
                "all": {("k_" + k): {"ms": round(v["ms"], 4), "achieved": round(v["gbs"], 2), "frac": round(v["gbs"] / peak, 4)} for k, v in own.items()},
                "all_note": "stage times are CUDA events on the launching stream inside the resident step; k_prefix shares the GPU with k_target (own stream)"}
    ms_per_step = 1000.0 * t_res / K
    e2e_ms = 1000.0 * t_e2e / K
    line = {
        "metric": METRIC, "value": n_res / t_res, "unit": "fragments/s", "n_gpus": world,
        "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64+u64", "data": "synthetic", "verified": verified, "verification": verify_note,
        "config": {"workload": "C3 synthetic human-transcriptome-scale: %d transcripts / %d nt / %d molecules per GPU, %d requested "
                               "fragments in total, 15 cycles, after_prim_double" % (n_tr, int(weak.bases_np.size), weak_stats["n_molecules"], weak.n_frags_global),
                   "flags": " ".join(C3_FLAGS), "scale": opts.scale,
                   "sharding": "contiguous transcripts per rank; exchange inside the library (NCCL on the handle's stream): all-gather of "
                               "[per-length totals | partial target histogram], grouped send/recv of the drawn copy ranks" if pool.lib_comm else
                               "contiguous transcripts per rank; torch.distributed all-gather + all-to-all",
                   "cpu_binding_rank0": (sorted(cpus)[0], sorted(cpus)[-1], len(cpus)) if cpus else None,
                   "l2": "inputs (0.45 GB bases + %.1f GB prefix sums per GPU) exceed the 126 MB L2; no explicit flush" % ((1 if mirrored else 2) * 4.125 * nt_total / 1e9),
                   "reverse_profile": "mirrored forward sums" if mirrored else "explicit arrays",
                   "e2e_path": ("chunk pipeline: every 32 MB chunk is ingested, fragmented, deduplicated and amplified on four streams while later "
                                "chunks are copied, one lagged host read per chunk (%d chunks in the last step)" % e2e_groups) if e2e_eager else "monolithic",
                   "species_per_gpu": weak_stats["n_species"], "fragments_after_fragmentation_per_gpu": weak_stats["n_fragments"],
                   "pool_total_rank0": weak_stats["pool_total"]},
        "e2e": {"value": n_e2e / t_e2e, "unit": "fragments/s", "h2d_bytes_per_step": int(weak.h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": e2e_ms, "phases_ms_rank0": {k: round(v / K, 3) for k, v in weak.phase.items()},
                "records": "packed: u32 count per transcript + u32 start + u16 length|strand per record (6 B/record; ids and transcript "
                           "indices are positional)",
                "copy_ceiling_ms": ceiling_ms,
                "frac_of_copy_ceiling": (ceiling_ms / e2e_ms) if ceiling_ms else None},
        "e2e_pageable": e2e_pageable,
        "e2e_packed_in": e2e_packed_in,
        "e2e_text_in": e2e_text_in,
        "e2e_dropin": e2e_dropin,
        "strong": strong,
        "stage_rates": {  # per-stage rates of the device-resident step
            "prefix_nt_per_s": nt_total / (stage_ms["prefix"] * 1e-3) if stage_ms.get("prefix", 0) > 0 else None,
            "fragmentation_molecules_per_s": float(weak_stats["n_molecules"]) / ((stage_ms["fragment"] + stage_ms.get("intervals", 0.0)) * 1e-3),
            "pcr_species_per_s": float(weak_stats["n_species"]) / (stage_ms["pcr"] * 1e-3),
            "selection_fragments_per_s": float(weak_stats["n_sampled"]) / ((stage_ms["select"] + stage_ms["expand"]) * 1e-3),
        },
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "kernels_ms_per_step": {k: round(v["ms"], 4) for k, v in kernels.items()},
        "kernels_gbs": {k: round(v["gbs"], 2) for k, v in own.items()},
    }
    if world == 1 and not opts.no_cpu_baseline:
        # the reference's release binary on this box's host cores, bounded sample (reported baseline, not the target) - and
        # the GPU path on exactly that sample through the same boundary (FASTA text in, FASTA text out): like for like
        try:
            ref = run_reference(REF_SCALE * opts.scale, 1, 0, REF_TIME_BUDGET_S)
            line["cpu_baseline"] = {"value": ref["value"], "unit": "fragments/s", "cores": ref["cores"], "kind": ref["kind"],
                                    "sample": ref["sample"], "stage_rates": ref.get("stages")}
            n_s, nf_s, sb_, so_, sl_ = reference_sample(REF_SCALE * opts.scale)
            raw = torch.from_numpy(synth.fasta_text(sb_, so_, sl_, width=0).copy()).pin_memory()
            names = [b"T%07d" % i for i in range(n_s)]
            names_u8 = np.frombuffer(b"".join(names), dtype=np.uint8).copy()
            names_off = np.arange(n_s + 1, dtype=np.uint64) * 8
            tbuf = None
            ts = []
            for s in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                pool.reset().configure(make_args(nf_s, 900 + s))
                h.set_first_index(0)
                h.add_fasta(raw, 1.0)
                h.build_library()
                h.sample()
                if tbuf is None:
                    tbuf = torch.empty(int(h.fasta_into(names_u8, names_off, None) * 1.1) + (1 << 20), dtype=torch.uint8).pin_memory()
                h.fasta_into(names_u8, names_off, tbuf)
                ts.append(time.perf_counter() - t0)
            n_same = int(h.stats().n_sampled)
            t_same = float(np.mean(ts[1:]))
            line["same_config"] = {"workload": ref["sample"], "gpu_fragments_per_s": n_same / t_same, "gpu_ms": 1000.0 * t_same,
                                   "reference_fragments_per_s": ref["value"], "reference_ms": ref["ms_per_step"],
                                   "ratio": (n_same / t_same) / ref["value"],
                                   "note": "both arms on the same sample through the same boundary (FASTA text in -> FASTA text out); "
                                           "the reference also reads its file and writes its JSON report"}
        except Exception as e:  # the baseline must never break the GPU line
            line["cpu_baseline"] = {"value": None, "unit": "fragments/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": "failed: %s" % e}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
