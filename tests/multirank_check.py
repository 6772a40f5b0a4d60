"""Multi-rank parity on real GPUs (run under torchrun, one rank per GPU; tests/test_gpu_multirank.py launches it):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 \
        tests/multirank_check.py [out.json]

Every rank owns a contiguous shard of a C3-shaped set (shard.contiguous_partition), fragments and amplifies it, and the
ranks select together - (a) through the library's own NCCL exchange (rlsim_sample_distributed: all-gather of totals +
partial target histograms, grouped send/recv of the drawn copy ranks on the handle's stream) and (b) through the
torch.distributed exchange (shard.RecordExchange).  Rank 0 gathers the ranks' records and compares their concatenation,
record for record, with the result of ONE GPU over the whole set (same seeds); the report histograms must agree too.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def digest(rec):
    return hashlib.sha256(b"".join(np.ascontiguousarray(rec[k]).astype("<u8").tobytes() for k in ("tr", "start", "end", "minus", "id"))).hexdigest()


def main():
    import torch
    import torch.distributed as dist
    import util
    from rlsim_b200 import _native, shard, synth
    from rlsim_b200.pool import GpuPool, GpuSampler

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    results = {}
    cases = [
        ("c3_shaped", 6000, 3.0e5, ["-n", "300000", "-si", "11", "-d", "0.9:sn:(600,50,4,300,2000) + 0.1:n:(700,1,600,2000)",
                                    "-f", "after_prim_double", "-c", "15", "-eg", "(1.0,0.5,0.8)", "-el", "(0.1,0.7,1.0)"]),
        ("exhaustion", 400, 2.0e4, ["-n", "120000", "-c", "2", "-si", "5", "-d", "1:n:(150,3,140,160)"]),  # missing + complement classes
        ("raw_target", 300, 2.0e4, ["-j", util.RAW_PARAMS, "-n", "20000", "-si", "6"]),
    ]
    for name, n_tr, n_mol, argv in cases:
        bases, off, levels = synth.transcriptome(n_tr, seed=1234, total_molecules=n_mol)
        args = util.parse(argv)
        lens = (off[1:] - off[:-1]).astype(np.int64)
        b = shard.contiguous_partition(shard.work_estimate(lens, levels), world)
        a0, a1 = b[rank], b[rank + 1]
        for mode in ("library", "torch"):
            if mode == "torch":
                os.environ["RLSIM_TORCH_EXCHANGE"] = "1"
            else:
                os.environ.pop("RLSIM_TORCH_EXCHANGE", None)
            pool = GpuPool(local, rank, world, dev)
            assert pool.lib_comm == (mode == "library"), (mode, pool.lib_comm)
            pool.configure(args)
            sub_off = (off[a0:a1 + 1] - off[a0]).astype(np.uint64)
            pool.init_transcripts_arrays(bases[int(off[a0]):int(off[a1])], sub_off, levels[a0:a1].copy(), a1 - a0, first=a0)
            rec = GpuSampler().sample_fragments(pool)
            hists = {k: pool.handle.hist(k, 4096).tolist() for k in (_native.H_TARGET, _native.H_SAMPLED, _native.H_MISSING, _native.H_AFTER_SAMPLING)}
            gathered = [None] * world
            dist.all_gather_object(gathered, {k: np.asarray(v) for k, v in rec.items()})
            if rank == 0:
                cat = {k: np.concatenate([g[k] for g in gathered]) for k in ("tr", "start", "end", "minus", "id")}
                if name not in results:  # one GPU over the whole set
                    h = _native.Handle(local)
                    h.set_model(args.to_params())
                    if args.raw_len_probs is not None:
                        h.set_raw_target(args.raw_len_probs.lengths, args.raw_len_probs.probs)
                    h.sample_target(args.req_frags)
                    h.add_transcripts_raw(bases, off, levels, n_tr)
                    h.build_library()
                    h.sample()
                    one = h.sampled()
                    results[name] = {"one_gpu": digest(one), "n": int(len(one["tr"])),
                                     "hists": {k: h.hist(k, 4096).tolist() for k in hists}, "missing": int(h.hist(_native.H_MISSING, 4096).sum())}
                    results[name]["_one"] = one
                one = results[name]["_one"]
                for k in ("tr", "start", "end", "minus", "id"):
                    assert np.array_equal(cat[k], one[k]), (name, mode, k)
                assert hists == results[name]["hists"], (name, mode, "report histograms")
                results[name][mode] = digest(cat)
                results[name]["per_rank_" + mode] = [int(len(g["tr"])) for g in gathered]
            pool.handle.close()
            dist.barrier()
    if rank == 0:
        for v in results.values():
            v.pop("_one", None)
            v.pop("hists", None)
            assert v["library"] == v["one_gpu"] == v["torch"]
        out = {"world": world, "cases": results, "ok": True}
        print(json.dumps(out))
        if len(sys.argv) > 1:
            json.dump(out, open(sys.argv[1], "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
