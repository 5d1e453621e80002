"""world_size-2 gloo test of the multi-rank host path: LPT sharding + result gather.
The per-rank engine call is replaced by the CPU oracle (tests may use it; the product never does)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import oracle
    from bialign_b200 import workloads
    from bialign_b200.batch import gather_scores, lpt_shards, pair_cost

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = dict(workloads.PROTEIN_PARAMS, max_shift=1)
    res, cls, off, pa, pb = workloads.protein_pairs(10, lo=5, hi=30, seed=11)
    lens = np.diff(off)
    shards = lpt_shards(pair_cost(lens[pa], lens[pb], 1), world)
    mine = shards[rank]
    scores = []
    for p in mine:
        a, sa = workloads.decode_protein(res, cls, off, int(pa[p]))
        b, sb = workloads.decode_protein(res, cls, off, int(pb[p]))
        scores.append(oracle.run(a, b, sa, sb, params, mode="codes")["score"])
    full = gather_scores(mine, scores, len(pa))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full)
    np.save(os.path.join(out_dir, f"mine{rank}.npy"), mine)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    m0, m1 = np.load(tmp_path / "mine0.npy"), np.load(tmp_path / "mine1.npy")
    assert (r0 == r1).all() and (r0 != 0).any()
    assert sorted(np.concatenate([m0, m1]).tolist()) == list(range(10))  # disjoint cover
    # single-process answer
    sys.path.insert(0, HERE)
    import oracle
    from bialign_b200 import workloads

    params = dict(workloads.PROTEIN_PARAMS, max_shift=1)
    res, cls, off, pa, pb = workloads.protein_pairs(10, lo=5, hi=30, seed=11)
    for p in range(10):
        a, sa = workloads.decode_protein(res, cls, off, int(pa[p]))
        b, sb = workloads.decode_protein(res, cls, off, int(pb[p]))
        assert oracle.run(a, b, sa, sb, params, mode="codes")["score"] == r0[p]
