"""Multi-GPU sharding of the pool.

Transcripts are independent through fragmentation and PCR (pool.go:101-124), so each rank owns a CONTIGUOUS range
of the input transcripts, balanced by the work estimate expression x length.  Contiguity is what makes the result
independent of the number of GPUs: every random stream is keyed by the global transcript index, and the per-length
rank space of the without-replacement draw is "species in (transcript, start, end) order", which the concatenation
of the ranks' shards reproduces exactly.

The only data that crosses NVLink: each rank's per-length pool totals T_r[l] (u64[high+1], <= 14 KB), one
all-gather (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def contiguous_partition(work, n_ranks):
    """Boundaries b[0..n_ranks] of contiguous ranges with ~equal total work (prefix quantiles)."""
    work = np.asarray(work, dtype=np.float64)
    n = len(work)
    csum = np.concatenate([[0.0], np.cumsum(work)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, n_ranks):
        target = total * r / n_ranks
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return bounds


def work_estimate(lengths, levels):
    return np.asarray(lengths, dtype=np.float64) * np.asarray(levels, dtype=np.float64)


def allgather_totals(local_totals, group=None, device=None):
    """-> (global_totals, rank_base) for this rank.  Single process: (local, zeros)."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local_totals, dtype=np.uint64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local.copy(), np.zeros_like(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = torch.from_numpy(local.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t, group=group)
    allt = np.stack([g.cpu().numpy().view(np.uint64) for g in gathered])  # [world, n_len]
    total = np.zeros_like(local)
    for r in range(world):
        nt = total + allt[r]
        if np.any(nt < total):
            raise OverflowError("Integer overflow when updating total fragment count!")  # fragstats.go:88-94
        total = nt
    base = allt[:rank].sum(axis=0, dtype=np.uint64) if rank > 0 else np.zeros_like(local)
    return total, base


def allreduce_hist(h, group=None, device=None):
    """Sum a per-rank u64 histogram over the ranks (report panels)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return h
    t = torch.from_numpy(np.ascontiguousarray(h, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, group=group)
    return t.cpu().numpy().view(np.uint64)
