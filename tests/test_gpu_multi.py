"""Multi-GPU parity on hardware: the in-library multi-device handle and the torchrun (NCCL) path must return exactly
what one GPU returns -- scores, traces, completeness flags, in the caller's pair order."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

pytestmark = pytest.mark.gpu

PARAMS = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
              shift_cost=-150, max_shift=2)


def _workload(npairs=600, lo=20, hi=260, seed=31):
    from bialign_b200 import workloads

    return workloads.protein_pairs(npairs, lo=lo, hi=hi, seed=seed)


def _single(res, cls, off, pa, pb, want_trace=True):
    from bialign_b200.batch import BatchAligner

    return BatchAligner(device=0, **PARAMS).align_encoded(res, cls, off, pa, pb, want_trace=want_trace)


def _same(a, b):
    assert (a[0] == b[0]).all()
    assert (a[2] == b[2]).all() and (a[1] == b[1]).all() and (a[3] == b[3]).all()


def test_multi_handle_on_one_device_equals_plain_engine():
    """ba_engine_create_multi with a single device: the sharding / merge code path on a one-GPU box."""
    from bialign_b200.batch import BatchAligner

    res, cls, off, pa, pb = _workload(200)
    want = _single(res, cls, off, pa, pb)
    al = BatchAligner(devices=[0], **PARAMS)
    assert al.engine.n_devices == 1
    got = al.align_encoded(res, cls, off, pa, pb, want_trace=True)
    _same(want, got)
    st = al.engine.stats()
    assert st["pairs"] == 200 and st["cell_states"] > 0
    assert (al.align_encoded(res, cls, off, pa, pb, want_trace=False) == want[0]).all()
    ev = al.engine.debug_end_values(17)
    assert ev.max() == want[0][17]


def test_multi_handle_over_all_gpus_equals_one_gpu():
    import torch
    from bialign_b200.batch import BatchAligner

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    res, cls, off, pa, pb = _workload(5000)  # > 4096 pairs: the snake deal; below: exact LPT
    want = _single(res, cls, off, pa, pb)
    al = BatchAligner(devices="all", **PARAMS)
    assert al.engine.n_devices == torch.cuda.device_count()
    _same(want, al.align_encoded(res, cls, off, pa, pb, want_trace=True))
    _same(tuple(x[:300] if i == 0 else x for i, x in enumerate(_single(res, cls, off, pa[:300], pb[:300]))),
          al.align_encoded(res, cls, off, pa[:300], pb[:300], want_trace=True))


def _rank_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from bialign_b200.batch import BatchAligner, gather_scores, gather_traces

    res, cls, off, pa, pb = _workload(900)
    al = BatchAligner(device=rank, **PARAMS)
    mine, (scores, cols, toff, complete) = al.align_sharded(res, cls, off, pa, pb, rank, world, want_trace=True)
    dev = torch.device("cuda", rank)
    full = gather_scores(mine, scores, len(pa), device=dev)
    gcols, goff, gcomp = gather_traces(mine, cols, toff, complete, len(pa), device=dev)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), scores=full, cols=gcols, off=goff, comp=gcomp)
    dist.barrier()
    dist.destroy_process_group()


def test_torchrun_path_over_two_gpus_equals_one_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_rank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = np.load(tmp_path / "gathered.npz")
    res, cls, off, pa, pb = _workload(900)
    want = _single(res, cls, off, pa, pb)
    _same(want, (g["scores"], g["cols"], g["off"], g["comp"]))
