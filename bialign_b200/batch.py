"""Batch front-end: many independent pairs over one sequence table, sharded across GPUs.

Pairs are independent, so the multi-GPU path has no data-path collective: rank r aligns its LPT
share of the pair list on its own GPU and the scores are gathered afterwards."""
import heapq

import numpy as np

from . import encoding
from ._capi import get_engine


def pair_cost(len_a, len_b, max_shift):
    """Work of one pair in band cells: (n+1)(m+1)(2s+1)^2 (upper bound of C(n,m,s))."""
    w = 2 * max_shift + 1
    return (np.asarray(len_a, dtype=np.int64) + 1) * (np.asarray(len_b, dtype=np.int64) + 1) * w * w


def band_cells(n, m, s):
    """Exact number of band-valid cells C(n,m,s): k in [max(0,i-s), min(n,i+s)], l likewise."""
    n = np.asarray(n, dtype=np.int64)
    m = np.asarray(m, dtype=np.int64)

    def one(length):
        tot = np.zeros_like(length)
        for d in range(-s, s + 1):
            tot = tot + np.maximum(length + 1 - abs(d), 0)
        return tot

    return one(n) * one(m)


def cell_states(n, m, s, affine=True):
    return (9 if affine else 1) * band_cells(n, m, s)


def lpt_shards(costs, world_size):
    """Longest-processing-time-first partition: returns a list of index arrays, one per rank.
    Deterministic (ties broken by pair index), identical on every rank."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.lexsort((np.arange(costs.size), -costs))
    if world_size == 1:
        return [np.sort(order)]
    shards = [[] for _ in range(world_size)]
    if costs.size > 4096:
        # snake (boustrophedon) deal of the sorted list: within 1/world_size of LPT for large batches
        # and O(n); exact LPT below for small ones
        idx = np.arange(order.size)
        rnd, pos = idx // world_size, idx % world_size
        rank = np.where(rnd % 2 == 0, pos, world_size - 1 - pos)
        return [np.sort(order[rank == r]) for r in range(world_size)]
    heap = [(0, r) for r in range(world_size)]
    for p in order:
        load, r = heapq.heappop(heap)
        shards[r].append(int(p))
        heapq.heappush(heap, (load + int(costs[p]), r))
    return [np.array(sorted(sh), dtype=np.int64) for sh in shards]


class BatchAligner:
    """Scores (and optionally traces) for a list of pairs under one scoring model."""

    def __init__(self, type="Protein", simmatrix=None, structure_weight=400, gap_opening_cost=0, gap_cost=-200,
                 shift_cost=-250, max_shift=2, sequence_match_similarity=100, sequence_mismatch_similarity=0,
                 device=None, **_ignored):
        self.type = type
        self.max_shift = int(max_shift)
        self.params = dict(structure_weight=int(structure_weight), gap_opening_cost=int(gap_opening_cost),
                           gap_cost=int(gap_cost), shift_cost=int(shift_cost), max_shift=int(max_shift))
        if simmatrix:
            matrix = encoding.read_simmatrix(simmatrix)
            self.symbols, self.table, _ = encoding.simmatrix_table(matrix)
        else:
            # match/mismatch scoring (pyx:409-412): residues are raw bytes, re-coded densely per batch
            self.symbols = None
            self._match = (int(sequence_match_similarity), int(sequence_mismatch_similarity))
            self.table = encoding.match_table(*self._match, nsym=4)
        self._device = device
        self._engine = None

    @property
    def engine(self):
        if self._engine is None:
            self._engine = get_engine(self._device)
        return self._engine

    def encode(self, seqs, structs):
        """Strings -> (residues, classes, offsets) of the C ABI."""
        res, cls, off = [], [], [0]
        for sq, st in zip(seqs, structs):
            if len(sq) != len(st):
                raise ValueError("Provided structure and sequence must have the same length.")
            res.append(encoding.encode_residues(sq, self.symbols) if self.symbols else encoding.encode_bytes(sq))
            cls.append(encoding.rna_structure_classes(st) if self.type == "RNA" else encoding.encode_bytes(st))
            off.append(off[-1] + len(sq))
        cat = lambda xs: np.concatenate(xs).astype(np.uint8) if xs else np.zeros(0, dtype=np.uint8)  # noqa: E731
        res = cat(res)
        if not self.symbols:
            used, res = np.unique(res, return_inverse=True)
            res = res.astype(np.uint8)
            self.table = encoding.match_table(*self._match, nsym=max(len(used), 1))
        return res, cat(cls), np.array(off, dtype=np.int64)

    def configure(self):
        p = self.params
        self.engine.set_scoring(self.table, p["structure_weight"], p["gap_opening_cost"], p["gap_cost"],
                                p["shift_cost"], p["max_shift"])

    def align_encoded(self, res, cls, off, pair_a, pair_b, want_trace=False):
        """End-to-end on host arrays.  Returns scores, or (scores, cols, offsets, complete)."""
        self.configure()
        scores = self.engine.align_batch(res, cls, off, pair_a, pair_b, want_trace=want_trace)
        if not want_trace:
            return scores
        cols, offsets, complete = self.engine.fetch_traces()
        return scores, cols, offsets, complete

    def align(self, seqs, structs, pairs, want_trace=False):
        res, cls, off = self.encode(seqs, structs)
        pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
        return self.align_encoded(res, cls, off, pairs[:, 0], pairs[:, 1], want_trace)

    def align_sharded(self, res, cls, off, pair_a, pair_b, rank, world_size, want_trace=False):
        """This rank's LPT share of the pair list.  Returns (global pair indices, results...)."""
        lens = np.diff(off)
        costs = pair_cost(lens[np.asarray(pair_a)], lens[np.asarray(pair_b)], self.max_shift)
        mine = lpt_shards(costs, world_size)[rank]
        sres, scls, soff, spa, spb = compact_shard(res, cls, off, np.asarray(pair_a)[mine], np.asarray(pair_b)[mine])
        out = self.align_encoded(sres, scls, soff, spa, spb, want_trace)
        return (mine, out)


def trace_hex(cols, offsets, p):
    return "".join("%x" % c for c in cols[offsets[p]:offsets[p + 1]])


def compact_shard(res, cls, off, pair_a, pair_b):
    """Sequence table restricted to the sequences a shard's pairs use (so a rank uploads only its share).
    Returns (res, cls, off, pair_a, pair_b) with re-numbered sequence indices."""
    pair_a = np.asarray(pair_a, dtype=np.int64)
    pair_b = np.asarray(pair_b, dtype=np.int64)
    used, inv = np.unique(np.concatenate([pair_a, pair_b]), return_inverse=True)
    lens = (off[used + 1] - off[used]).astype(np.int64)
    new_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.arange(int(new_off[-1])) - np.repeat(new_off[:-1], lens) + np.repeat(off[used], lens)
    return (np.ascontiguousarray(res[idx]), np.ascontiguousarray(cls[idx]), new_off,
            inv[:len(pair_a)].astype(np.int32), inv[len(pair_a):].astype(np.int32))


def gather_scores(mine, scores, n_total, device=None):
    """All ranks' scores in caller order.  `mine` = this rank's pair indices (from lpt_shards),
    `scores` = their scores.  Uses the default torch.distributed group (NCCL on GPUs, gloo in CPU
    tests); with a single process it is a plain scatter.  This is the only cross-rank step of the
    batch path -- a result gather, not a data-path collective."""
    import torch
    import torch.distributed as dist

    full = torch.zeros(n_total, dtype=torch.int64, device=device)
    full[torch.as_tensor(np.asarray(mine), dtype=torch.int64, device=device)] = torch.as_tensor(
        np.asarray(scores, dtype=np.int64), device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM)  # shards are disjoint, so SUM == concatenation
    return full.cpu().numpy()
