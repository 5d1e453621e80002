#!/usr/bin/env python3
"""BASELINE config 4 at full size on one GPU: 1 000 000 RNA pairs of length 120, score only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bialign_b200 import workloads
from bialign_b200.batch import BatchAligner
t0 = time.perf_counter()
res, cls, off, pa, pb = workloads.rna_pairs(1000000, seed=4)
t1 = time.perf_counter()
al = BatchAligner(max_shift=2, **workloads.RNA_PARAMS)
scores = al.align_encoded(res, cls, off, pa, pb, want_trace=False)
t2 = time.perf_counter()
scores2 = al.align_encoded(res, cls, off, pa, pb, want_trace=False)
t3 = time.perf_counter()
st = al.engine.stats()
print("generate %.1f s; first call %.2f s; second call (end to end, host buffers) %.2f s; device %.1f ms; %.1f GCUPS e2e; kind %d warps %d; checksum %d"
      % (t1 - t0, t2 - t1, t3 - t2, st["total_ms"], st["cell_states"] / (t3 - t2) / 1e9, st["kernel_kind"], st["warps_per_cta"], int(scores.sum())))
assert (scores == scores2).all()
