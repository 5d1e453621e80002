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
    """Scores (and optionally traces) for a list of pairs under one scoring model.

    Scoring arguments are the reference's (BiAligner's **params, pyx:179-193); anything else raises TypeError.
    `device` = one CUDA ordinal (default: LOCAL_RANK); `devices` = a list of ordinals or "all": the pair list is
    then sharded over those GPUs inside the library (one host thread per GPU, no collective).
    Engines are shared per device within the process, so tuning options belong to the aligner (`set_option`) and are
    applied -- all others reset to automatic -- each time it configures the engine; one thread at a time per engine."""

    def __init__(self, type="Protein", simmatrix=None, structure_weight=400, gap_opening_cost=0, gap_cost=-200,
                 shift_cost=-250, max_shift=2, sequence_match_similarity=100, sequence_mismatch_similarity=0,
                 device=None, devices=None, nameA=None, nameB=None, outmode=None, nodescription=None):
        # (nameA, nameB, outmode, nodescription: presentation parameters of the reference's param dict, unused here)
        self.type = type
        self.max_shift = int(max_shift)
        self.params = dict(structure_weight=int(structure_weight), gap_opening_cost=int(gap_opening_cost),
                           gap_cost=int(gap_cost), shift_cost=int(shift_cost), max_shift=int(max_shift))
        self.known = None
        if simmatrix:
            matrix = encoding.read_simmatrix(simmatrix)
            self.symbols, self.table, self.known = encoding.simmatrix_table(matrix)
        else:
            # match/mismatch scoring (pyx:409-412): residues are raw bytes, re-coded densely per batch
            self.symbols = None
            self._match = (int(sequence_match_similarity), int(sequence_mismatch_similarity))
            self.table = encoding.match_table(*self._match, nsym=4)
        self._device = device
        self._devices = devices
        self._engine = None
        self.options = {}

    @property
    def engine(self):
        if self._engine is None:
            self._engine = get_engine(self._device, self._devices)
        return self._engine

    def set_option(self, key, value):
        """Tuning option of this aligner (see ba_set_option); validated by the library right away."""
        self.engine.set_option(key, value)
        self.options[key] = int(value)

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

    def check_known(self, res, off, pair_a, pair_b):
        """KeyError when a pair needs a residue combination the similarity matrix does not define -- what the reference's
        dict-of-dicts lookup raises (pyx:407).  Only matrices with holes (sparse / asymmetric files) need the check."""
        if self.known is None or self.known.all():
            return
        off = np.asarray(off)
        nsym = self.known.shape[0]
        present = np.zeros((len(off) - 1, nsym), dtype=bool)  # which symbols each sequence contains
        seq_of = np.repeat(np.arange(len(off) - 1), np.diff(off))
        present[seq_of, np.asarray(res)[off[0]:off[-1]]] = True
        pa, pb = np.asarray(pair_a), np.asarray(pair_b)
        for a, b in set(zip(pa.tolist(), pb.tolist())):
            bad = present[a][:, None] & present[b][None, :] & ~self.known
            if bad.any():
                i, j = np.argwhere(bad)[0]
                raise KeyError(self.symbols[j])  # the inner key of simmatrix[a][b]

    def configure(self):
        eng, p = self.engine, self.params
        eng.apply_options(self.options)  # options of whoever used the shared engine before do not leak in
        eng.set_scoring(self.table, p["structure_weight"], p["gap_opening_cost"], p["gap_cost"],
                        p["shift_cost"], p["max_shift"])

    def align_encoded(self, res, cls, off, pair_a, pair_b, want_trace=False, mu2=None):
        """End-to-end on host arrays.  Returns scores, or (scores, cols, offsets, complete).
        `mu2`: optional list of per-pair int matrices len(A) x len(B) replacing the class-equality structure similarity
        (probabilistic RNA profiles, pyx:414-423); such batches run on the general level kernel."""
        self.check_known(res, off, pair_a, pair_b)
        self.configure()
        if mu2 is not None:
            eng = self.engine
            eng.load_sequences(res, cls, off)
            eng.load_pairs(pair_a, pair_b)
            eng.set_pair_mu2(mu2)
            eng.run(want_trace=want_trace)
            scores = eng.fetch_scores()
            if not want_trace:
                return scores
            cols, offsets, complete = eng.fetch_traces()
            return scores, cols, offsets, complete
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


def gather_traces(mine, cols, offsets, complete, n_total, device=None, dst=None):
    """All ranks' traces in caller order: returns (cols uint8, offsets int64[n_total+1], complete uint8[n_total]).
    `mine` = this rank's pair indices, (cols, offsets, complete) = what its engine returned for them.  Trace lengths
    and flags travel like the scores (disjoint scatters, SUM); the columns are placed at their global positions in a flat
    byte buffer that is then SUM-reduced -- supports are disjoint, so the sum is the concatenation.  With `dst` only
    that rank receives the columns (torch.distributed.reduce); the others get an empty cols array."""
    import torch
    import torch.distributed as dist

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    mine = np.asarray(mine, dtype=np.int64)
    offsets = np.asarray(offsets, dtype=np.int64)
    lens = np.diff(offsets)
    meta = torch.zeros((2, n_total), dtype=torch.int64, device=device)
    idx = torch.as_tensor(mine, device=device)
    meta[0, idx] = torch.as_tensor(lens, device=device)
    meta[1, idx] = torch.as_tensor(np.asarray(complete, dtype=np.int64), device=device)
    if multi:
        dist.all_reduce(meta, op=dist.ReduceOp.SUM)
    meta_h = meta.cpu().numpy()
    goff = np.concatenate([[0], np.cumsum(meta_h[0])]).astype(np.int64)
    total_local = int(offsets[-1] - offsets[0])
    flat = torch.zeros(int(goff[-1]), dtype=torch.uint8, device=device)
    if total_local:
        # byte k of local pair p goes to goff[mine[p]] + k: per-pair shifts expanded on the device (no per-byte host index)
        shift = torch.as_tensor(goff[mine] - (offsets[:-1] - offsets[0]), device=device)
        place = torch.repeat_interleave(shift, torch.as_tensor(lens, device=device)) + torch.arange(total_local, device=device)
        flat[place] = torch.as_tensor(np.ascontiguousarray(cols[offsets[0]:offsets[-1]]), device=device)
    if multi:
        if dst is None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        else:
            dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM)
            if dist.get_rank() != dst:
                flat = flat[:0]
    return flat.cpu().numpy(), goff, meta_h[1].astype(np.uint8)

