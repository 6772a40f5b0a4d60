"""rlsim driver over the GPU pool: `python -m rlsim_b200 [flags] transcripts.fas > fragments.fas`.

Mirror of /root/reference/src/main.go:41-162: same flags, same FASTA in/out, same JSON report; the Pooler and
Sampler are the GPU-backed ones.  Under torchrun each rank handles a contiguous shard of the transcripts and
writes its own records; rank 0 writes the report.
"""
import os
import sys
import time

import numpy as np

from . import _native, effs
from .args import ArgError, CmdArgs, ExitRequest
from .fasta import load_transcripts, read_input
from .mixture import mixture_string
from .pool import FragStats, GpuPool, GpuSampler
from .report import Report


class Log:  # logging.go:49-122
    def __init__(self, verbose):
        self.verbose = verbose

    def _stamp(self):
        t = time.time()
        lt = time.localtime(t)
        return "%02d:%02d:%02d.%06d" % (lt.tm_hour, lt.tm_min, lt.tm_sec, int((t % 1) * 1e6))

    def println(self, msg):
        sys.stderr.write("%s %s\n" % (self._stamp(), msg))

    def v(self, msg):
        if self.verbose:
            self.println(msg)

    def fatal(self, msg):
        self.println(msg)
        sys.exit(1)


def run(argv, out=None):
    log = Log(True)
    try:
        args = CmdArgs.parse(argv)
    except ExitRequest as e:
        sys.stderr.write(e.text)
        sys.exit(e.code)
    except ArgError as e:
        log.fatal(str(e))
    log = Log(args.verbose)
    for w in args.warnings:
        log.println(w)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist_device = None
    if world > 1:
        import torch
        import torch.distributed as dist
        from . import shard
        shard.bind_to_gpu_numa_node(local_rank)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
        dist_device = torch.device("cuda", local_rank)
    log.v("Starting up rlsim using maximum %d cores." % args.max_procs)
    if args.init_seed == 0:
        args.init_seed = int(time.time())  # main.go:76-79
        if world > 1:  # every rank must key its streams with the logged seed: rank 0's clock decides
            box = [args.init_seed]
            dist.broadcast_object_list(box, src=0)
            args.init_seed = int(box[0])
    log.v("Initial random seed: %d" % args.init_seed)
    if args.raw_params_file:
        log.v("Using raw parameter file: %s" % args.raw_params_file)
    report = Report(args.report_file) if rank == 0 else None
    stats = FragStats()
    try:
        pool = GpuPool(local_rank, rank, world, dist_device).configure(args)
        pool.finish_target()  # the report wants the whole target histogram now (multi-GPU: summed over the ranks)
        h = pool.handle
        st = h.stats()
        log.v("Number of requested fragments: %d" % args.req_frags)
        if args.raw_len_probs is None:
            log.v("Target fragment length distribution components:")
            for line in mixture_string(args.target_mix).split("\n"):
                log.v(line)
        log.v("Finished sampling target lengths.")
        if report:
            report.report_map(h.hist_dict(_native.H_TARGET, int(st.high) + 2), "Length", "Count", "Target lenghts", "bar")
        log.v('Fragmentation method: "%s" with parameter %d.' % (args.frag_method.name, args.frag_method.param))
        log.v("Number of PCR cycles: %d" % args.nr_cycles)
        if report and not args.fixed_eff > 0.0:  # thermocycler.go:195-221
            xs, ys = effs.gc_panel(pool.params)
            report.report_float(xs, ys, "GC content (%)", "Efficiency", "GC efficiency function", "line")
            report.report_map(effs.len_panel(pool.params, st.low, st.high), "Length", "Efficiency",
                              "Length efficiency function", "line")
        log.v("Poly(A) tail distribution components:")
        for line in mixture_string(args.polya_param).split("\n"):
            log.v(line)
        levels = None
        data = read_input(args.input_files)
        log.v("Fragmenting transcripts and amplifying fragments:")
        if pool.world == 1 and not os.environ.get("RLSIM_HOST_FASTA"):
            # device-side FASTA reader (fasta.go:113-212 on the GPU); unusual texts fall back to the host reader
            try:
                levels = pool.init_transcripts_fasta(data, args.expr_mul, stats, log.println)
            except _native.RlsimError as e:
                if e.code != _native.E_UNSUPPORTED:
                    raise
                pool.reset()
        if levels is None:
            names, seqs, levels = load_transcripts(args.input_files, args.expr_mul, log.println, data=data)
            pool.init_transcripts(names, seqs, levels, stats)
        log.v("Initialized %d transcripts." % sum(1 for lv in levels if lv > 0))
        frags = GpuSampler(args.strand_bias).sample_fragments(pool, stats)
        fasta = h.fasta(pool.names)
        (out or sys.stdout.buffer).write(fasta)
        log.v("Total number of fragments in the pool: %g" % float(stats.total_frags))
        log.v("Sampling ratio: %g" % (stats.nr_sampled / float(stats.total_frags) if stats.total_frags else float("nan")))
        log.v("Sampled %d fragments." % stats.nr_sampled)
        log.v("Missing fragments: %d" % (args.req_frags - stats.nr_sampled))
    except (_native.RlsimError, OverflowError, ImportError, ValueError, IndexError, KeyError) as e:
        # the reference's only error policy is a fatal log line and exit 1 (logging.go:67-73): "Illegal fasta record"
        # (fasta.go:183-185) and malformed raw-parameter files (parse_raw.go) included - json.JSONDecodeError is a ValueError
        log.fatal(str(e))
    if report:  # fragstats.go:120-141
        report.report_map(stats.after_frag, "Length", "Count", "Fragdist after fragmentation", "bar")
        report.report_map(stats.after_pcr, "Length", "Count", "Fragdist after PCR", "bar")
        report.report_map(stats.after_sampling, "Length", "Count", "Fragdist after sampling", "bar")
        report.report_map(stats.missing, "Length", "Count", "Missing fragments", "bar")
        report.report_map(stats.sampled, "Length", "Count", "Sampled fragments", "bar")
        report.report_map(stats.polya_len, "Length", "Count", "Poly-A tail lengths", "bar")
        report.report_map(stats.tr_lengths, "Length", "Count", "Transcript lengths", "hist")
        report.report_map(stats.expr_levels, "Expression level", "Count", "Expression levels", "hist")
        ratio = stats.nr_sampled / float(stats.total_frags) if stats.total_frags else float("nan")
        report.report_float([0.0], [ratio], "", "Sampling ratio", "Sampling ratio", "bar")
        report.write_json()
    return frags


def main():
    run(sys.argv[1:])


if __name__ == "__main__":
    main()
