#!/usr/bin/env python3
"""Throughput of the rebased wide-range trace run vs the level kernel on gcd-1 scoring (development aid)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bialign_b200.batch import BatchAligner

rng = np.random.default_rng(5)
aa = "ARNDCQEGHILKMFPSTWYV"
def batch(n, lo, hi):
    seqs, structs, pairs = [], [], []
    for q in range(n):
        for _ in range(2):
            L = int(rng.integers(lo, hi + 1))
            seqs.append("".join(aa[i] for i in rng.integers(0, 20, L)))
            structs.append("".join("HEC"[i] * 6 for i in rng.integers(0, 3, L // 6 + 1))[:L])
        pairs.append((2 * q, 2 * q + 1))
    return seqs, structs, pairs
params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=333, gap_opening_cost=-157, gap_cost=-49, shift_cost=-151, max_shift=2)
std = dict(params, structure_weight=800, gap_opening_cost=-150, gap_cost=-50, shift_cost=-150)
for name, (n, lo, hi, opt, par) in {"3000 pairs 200-500 gcd1": (3000, 200, 500, -1, params),
                                    "800 pairs 2000 gcd1": (800, 1900, 2100, -1, params), "2 pairs 2000 gcd1": (2, 2000, 2000, -1, params),
                                    "800 pairs 2000 gcd50": (800, 1900, 2100, -1, std), "800 pairs 2000 gcd50 long=0": (800, 1900, 2100, -2, std),
                                    "40 pairs 2000 gcd50": (40, 1900, 2100, -1, std), "40 pairs 2000 gcd50 long=0": (40, 1900, 2100, -2, std)}.items():
    seqs, structs, pairs = batch(n, lo, hi)
    al = BatchAligner(**par)
    if opt == -2:
        al.set_option("long", 0)
    for rep in range(2):
        t0 = time.time(); out = al.align(seqs, structs, pairs, want_trace=True); t1 = time.time()
    st = al.engine.stats()
    print(name, "kind", st["kernel_kind"], "fallback", st["fallback_pairs"], "fill_ms %.2f tb_ms %.2f" % (st["fill_ms"], st["traceback_ms"]),
          "GCUPS %.1f" % (st["cell_states"] / st["fill_ms"] / 1e6), "wall %.3f s" % (t1 - t0), "complete", int(out[3].sum()), "/", len(pairs), flush=True)
