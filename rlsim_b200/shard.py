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


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU, so that pinned host buffers are first-touched on
    the NUMA node the GPU hangs off (with 8 ranks copying at once, remote-socket buffers halve the PCIe throughput).
    Best effort: returns the CPU set or None."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and all(v.strip().isdigit() for v in visible.split(",")) else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < n_cpu}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or None
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


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


def allgather_totals_matrix(local_totals, group=None, device=None):
    """-> [world, n_len] u64 matrix of every rank's per-length totals (row r = rank r)."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local_totals, dtype=np.uint64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local[None, :].copy()
    world = dist.get_world_size(group)
    t = torch.from_numpy(local.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t, group=group)
    return np.stack([g.cpu().numpy().view(np.uint64) for g in gathered])


class RecordExchange:
    """Size selection on N GPUs without redundant work.

    Length classes are partitioned over the ranks (class l is drawn by rank l % N); every drawn (or, for classes drawn
    more than half, excluded) copy rank is routed to the rank that owns the copy with ONE all-to-all over NVLink
    (16-byte records); the owner maps it to its species.  Same result as one GPU: the per-class streams do not depend on
    who evaluates them."""

    def __init__(self, device):
        self.device = device
        self.send = None

    def _buffer(self, torch, n):
        if self.send is None or self.send.shape[0] < n:
            self.send = torch.empty((int(n * 1.25) + 1024, 2), dtype=torch.int64, device=self.device)
        return self.send

    def run(self, handle, rank, world, expected_records, group=None, pool=None):
        import torch
        import torch.distributed as dist
        nb = getattr(pool, "_target_nb", 0) if pool is not None else 0
        if nb:  # the partial target histograms ride along with the per-length totals: one collective instead of two
            totals = handle.class_totals()
            both = allgather_totals_matrix(np.concatenate([totals, handle.hist(0, nb)]), group, self.device)
            all_tot = np.ascontiguousarray(both[:, :len(totals)])
            pool._sum_target(both[:, len(totals):].sum(axis=0, dtype=np.uint64))
        else:
            all_tot = allgather_totals_matrix(handle.class_totals(), group, self.device)
        buf = self._buffer(torch, expected_records)
        counts, needed = handle.select_classes(all_tot, rank, buf.data_ptr(), buf.shape[0])
        if needed > buf.shape[0]:
            buf = self._buffer(torch, needed)
            counts, needed = handle.select_classes(all_tot, rank, buf.data_ptr(), buf.shape[0])
        send_counts = torch.from_numpy(counts.astype(np.int64)).to(self.device)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        rc = [int(x) for x in recv_counts.cpu()]
        recv = torch.empty((sum(rc), 2), dtype=torch.int64, device=self.device)
        dist.all_to_all_single(recv, buf[:needed], output_split_sizes=rc, input_split_sizes=[int(c) for c in counts], group=group)
        if self.device is not None and torch.cuda.is_available():
            torch.cuda.synchronize(self.device)
        handle.apply_selected(recv.data_ptr() if recv.shape[0] else None, recv.shape[0])
        handle.finish_sample()
        return all_tot


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
