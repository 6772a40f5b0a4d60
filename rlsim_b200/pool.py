"""GpuPool / GpuSampler: the two batch seams of the reference, backed by librlsim_gpu.so.

    GpuPool.init_transcripts   = Pooler.InitTranscripts   (/root/reference/src/pool.go:83-135)
    GpuSampler.sample_fragments = Sampler.SampleFragments (/root/reference/src/sampler.go:108-154)

Everything between them (per-molecule fragmentation, per-species PCR, per-length pools, without-replacement
draws) runs on the device; this file only moves buffers and, with more than one rank, performs the single
all-gather of per-length pool totals.
"""
import numpy as np

from . import _native, shard
from .args import CmdArgs


class FragStats:
    """fragstats.go:45-56: the histograms the report needs, filled from the library's counters."""

    def __init__(self):
        self.expr_levels, self.tr_lengths = {}, {}
        self.after_frag = self.after_pcr = self.after_sampling = self.sampled = self.missing = self.polya_len = None
        self.total_frags = 0
        self.nr_sampled = 0

    def update_input(self, seq_len, level):  # input.go:114-117
        self.tr_lengths[seq_len] = self.tr_lengths.get(seq_len, 0) + level
        self.expr_levels[level & 0xffffffff] = self.expr_levels.get(level & 0xffffffff, 0) + 1


def _nz(h):
    return {int(i): int(h[i]) for i in np.nonzero(h)[0]}


class GpuPool:
    def __init__(self, device=0, rank=0, world=1, dist_device=None):
        self.handle = _native.Handle(device)
        self.rank, self.world, self.dist_device = rank, world, dist_device
        self.names = []
        self.first = 0
        self.lib_comm = False
        if world > 1:
            self._init_lib_comm()

    def _init_lib_comm(self):
        """Join the ranks' handles with the library's own NCCL communicator (rlsim_comm_init): afterwards the whole
        multi-GPU selection is one library call on the handle's stream.  The id travels through the process group that
        launched the ranks; without NCCL / a GPU process group (gloo-only CPU tests) the torch.distributed exchange of
        shard.RecordExchange remains."""
        import os
        try:
            import torch.distributed as dist
            if os.environ.get("RLSIM_TORCH_EXCHANGE") or not (dist.is_available() and dist.is_initialized()):
                return
            if self.dist_device is None:  # gloo (CPU) group: no GPUs to join
                return
            box = [_native.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            self.handle.comm_init(box[0], self.world, self.rank)
            self.lib_comm = True
        except _native.RlsimError:
            self.lib_comm = False

    def configure(self, args: CmdArgs):
        """NewFragmentor + NewTechne + NewTarget + NewLenSampler, main.go:88-127."""
        self.args = args
        self.params = args.to_params()
        h = self.handle
        h.set_model(self.params)
        if args.raw_len_probs is not None:
            h.set_raw_target(args.raw_len_probs.lengths, args.raw_len_probs.probs)
        self._target_nb = 0
        if self.world > 1:  # every rank draws 1/world of the target lengths; the histograms are summed (same total
            n = args.req_frags  # histogram on every rank: the stream is keyed by the global draw index)
            a, b = n * self.rank // self.world, n * (self.rank + 1) // self.world
            h.sample_target_range(a, b - a)  # asynchronous, on the library's own stream
            nb = int(max([c.high for c in args.target_mix] + (args.raw_len_probs.lengths if args.raw_len_probs else [0]))) + 1
            if args.raw_len_probs is not None and self.lib_comm:
                h.comm_sum_target()  # a raw target shapes the library (range, mean): summed over the ranks now
            elif args.raw_len_probs is not None:
                self._sum_target(shard.allreduce_hist(h.hist(_native.H_TARGET, nb), device=self.dist_device))
            else:
                self._target_nb = nb  # summed later, together with the per-length totals (finish_target / RecordExchange)
        else:
            h.sample_target(args.req_frags)
        return self

    def _sum_target(self, tot):
        nz = np.nonzero(tot)[0]
        self.handle.set_target_hist(nz.astype(np.uint32), tot[nz])
        self._target_nb = 0

    def finish_target(self):
        """Multi-GPU: sum the ranks' partial target histograms if that is still pending (no-op otherwise)."""
        if getattr(self, "_target_nb", 0) and self.lib_comm:
            self.handle.comm_sum_target()
            self._target_nb = 0
        elif getattr(self, "_target_nb", 0):
            self._sum_target(shard.allreduce_hist(self.handle.hist(_native.H_TARGET, self._target_nb), device=self.dist_device))

    def reset(self):
        """Start another library on the same GPU, keeping the device buffers."""
        self.handle.clear_transcripts()
        return self

    def init_transcripts_arrays(self, bases, offsets, levels, n, first=0):
        """pool.go:83-135 for a shard that is already one concatenated buffer (numpy arrays or pinned torch
        tensors): `bases` u8, `offsets` u64[n+1], `levels` u64[n]; `first` = global index of its first record."""
        h = self.handle
        self.first = first
        h.set_first_index(first)
        h.add_transcripts_raw(bases, offsets, levels, n)
        h.build_library()
        return self

    def init_transcripts(self, names, seqs, levels, stats: FragStats = None):
        """pool.go:83-135 for this rank's contiguous shard of the input."""
        h = self.handle
        n = len(seqs)
        if stats is not None:
            for s, lv in zip(seqs, levels):
                stats.update_input(len(s), lv)
        bounds = shard.contiguous_partition(shard.work_estimate([len(s) for s in seqs], levels), self.world)
        a, b = bounds[self.rank], bounds[self.rank + 1]
        self.first, self.names = a, names[a:b]
        h.set_first_index(a)
        h.add_transcripts(seqs[a:b], levels[a:b])
        return self._build(stats)

    def init_transcripts_fasta(self, data: bytes, expr_mul: float = 1.0, stats: FragStats = None, log=None):
        """pool.go:83-135 fed by the device-side FASTA reader (single GPU): `data` = raw text of the input files.
        -> levels of the added transcripts."""
        from .fasta import load_transcripts_device
        if self.world != 1:
            raise ValueError("the device FASTA reader serves the single-GPU path; shards are cut by the host reader")
        h = self.handle
        h.set_first_index(0)
        names, lens, levels = load_transcripts_device(h, data, expr_mul, log)
        if stats is not None:
            for ln, lv in zip(lens, levels):
                stats.update_input(ln, lv)
        self.first, self.names = 0, names
        self._build(stats)
        return levels

    def _build(self, stats):
        h = self.handle
        h.build_library()
        if stats is not None:
            st = h.stats()
            red = lambda x: shard.allreduce_hist(x, device=self.dist_device)
            stats.after_frag = _nz(red(h.hist(_native.H_AFTER_FRAG, int(st.high) + 1)))
            stats.polya_len = _nz(red(h.hist(_native.H_POLYA, int(st.polya_max) + 1)))
        return self


class GpuSampler:
    def __init__(self, strand_bias=0.5):
        self.strand_bias = strand_bias  # already part of the model block; kept for interface parity (sampler.go:59)

    def sample_fragments(self, pool: GpuPool, stats: FragStats = None):
        """sampler.go:108-154.  -> dict of arrays (tr, start, end, minus, id) for this rank's shard."""
        h = pool.handle
        if pool.world > 1 and pool.lib_comm:  # the whole exchange inside the library (NCCL on the handle's stream)
            h.sample_distributed()
            pool._target_nb = 0
            total = h.hist(_native.H_POOL_GLOBAL, int(h.stats().high) + 1)
        elif pool.world > 1:  # classes partitioned over the ranks, drawn copy ranks routed to their owners (all-to-all)
            if not hasattr(pool, "_exchange"):
                pool._exchange = shard.RecordExchange(pool.dist_device)
            all_tot = pool._exchange.run(h, pool.rank, pool.world, int(pool.args.req_frags) // pool.world + 4096, pool=pool)
            total = all_tot.sum(axis=0, dtype=np.uint64)
        else:
            total = h.class_totals()
            h.sample()
        if stats is not None:
            n = len(total)
            stats.after_pcr = _nz(total)
            stats.total_frags = int(total.sum(dtype=np.uint64)) if float(total.sum(dtype=np.float64)) < 1.8e19 else None
            if stats.total_frags is None:
                raise OverflowError("Integer overflow when updating total fragment count!")
            tn = max(n, int(h.stats().high) + 2)
            stats.sampled = {k: v for k, v in _nz(h.hist(_native.H_SAMPLED, tn)).items()}
            tgt = h.hist(_native.H_TARGET, tn)
            smp = h.hist(_native.H_SAMPLED, tn)
            mis = h.hist(_native.H_MISSING, tn)
            asm = h.hist(_native.H_AFTER_SAMPLING, tn)
            req = np.nonzero(tgt)[0]
            stats.sampled = {int(l): int(smp[l]) for l in req}    # sampler.go:95
            stats.missing = {int(l): int(mis[l]) for l in req}    # sampler.go:101
            stats.after_sampling = {int(l): int(asm[l]) for l in np.nonzero(total)[0]}  # fragstats.go:86,97-102
            stats.nr_sampled = int(smp.sum())
        return h.sampled()
