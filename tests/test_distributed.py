"""world_size-2 and -3 gloo tests of the multi-rank host path: LPT sharding + result gather.
The per-rank engine call is replaced by the CPU oracle (tests may use it; the product never does)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import oracle
    from bialign_b200 import workloads
    from bialign_b200.batch import gather_scores, gather_traces, lpt_shards, pair_cost

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = dict(workloads.PROTEIN_PARAMS, max_shift=1)
    res, cls, off, pa, pb = workloads.protein_pairs(10, lo=5, hi=30, seed=11)
    lens = np.diff(off)
    shards = lpt_shards(pair_cost(lens[pa], lens[pb], 1), world)
    mine = shards[rank]
    scores, cols, toff, comp = [], [], [0], []
    for p in mine:
        a, sa = workloads.decode_protein(res, cls, off, int(pa[p]))
        b, sb = workloads.decode_protein(res, cls, off, int(pb[p]))
        r = oracle.run(a, b, sa, sb, params, mode="codes")
        scores.append(r["score"])
        cols.extend(int(ch, 16) for ch in r["trace"])
        toff.append(len(cols))
        comp.append(1 if r["complete"] else 0)
    full = gather_scores(mine, scores, len(pa))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full)
    np.save(os.path.join(out_dir, f"mine{rank}.npy"), mine)
    # traces: to every rank, then to rank 1 only
    gc, go, gk = gather_traces(mine, np.array(cols, dtype=np.uint8), np.array(toff, dtype=np.int64), np.array(comp, dtype=np.uint8), len(pa))
    np.savez(os.path.join(out_dir, f"traces{rank}.npz"), cols=gc, off=go, comp=gk)
    dc, do, dk = gather_traces(mine, np.array(cols, dtype=np.uint8), np.array(toff, dtype=np.int64), np.array(comp, dtype=np.uint8), len(pa), dst=1)
    np.savez(os.path.join(out_dir, f"traces_dst{rank}.npz"), cols=dc, off=do, comp=dk)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])  # 10 pairs: an even and an uneven deal
def test_multi_rank_sharding_and_gather(tmp_path, world):
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    mines = [np.load(tmp_path / f"mine{r}.npy") for r in range(world)]
    r0 = ranks[0]
    assert all((r == r0).all() for r in ranks) and (r0 != 0).any()
    assert sorted(np.concatenate(mines).tolist()) == list(range(10))  # disjoint cover
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
    ts = [np.load(tmp_path / f"traces{r}.npz") for r in range(world)]
    ds = [np.load(tmp_path / f"traces_dst{r}.npz") for r in range(world)]
    t0 = ts[0]
    for k in ("cols", "off", "comp"):
        assert all((t[k] == t0[k]).all() for t in ts)
    # dst=1: only rank 1 holds the columns, every rank the offsets
    assert all(ds[r]["cols"].size == 0 for r in range(world) if r != 1)
    assert (ds[1]["cols"] == t0["cols"]).all() and all((d["off"] == t0["off"]).all() for d in ds)
    for p in range(10):
        a, sa = workloads.decode_protein(res, cls, off, int(pa[p]))
        b, sb = workloads.decode_protein(res, cls, off, int(pb[p]))
        r = oracle.run(a, b, sa, sb, params, mode="codes")
        got = "".join("%x" % c for c in t0["cols"][t0["off"][p]:t0["off"][p + 1]])
        assert got == r["trace"] and bool(t0["comp"][p]) == r["complete"]

