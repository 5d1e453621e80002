#!/usr/bin/env python3
"""Throughput of every BASELINE.json configuration shape on one GPU (device-resident inputs, CUDA-event times).
Not the contract bench (that is bench.py, config 3); this fills the per-config table of DESIGN.md."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bialign_b200 import workloads  # noqa: E402
from bialign_b200.batch import BatchAligner  # noqa: E402


def run(name, al, res, cls, off, pa, pb, want_trace, reps=3):
    eng = al.engine
    al.configure()
    eng.load_sequences(res, cls, off)
    eng.load_pairs(pa, pb)
    eng.run(want_trace=want_trace)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        eng.run(want_trace=want_trace)
        wall = time.perf_counter() - t0
        st = eng.stats()
        if best is None or st["total_ms"] < best["total_ms"]:
            best = dict(st, wall_ms=wall * 1e3)
    out = dict(config=name, pairs=int(len(pa)), want_trace=want_trace, cell_states=best["cell_states"],
               fill_ms=best["fill_ms"], traceback_ms=best["traceback_ms"], total_ms=best["total_ms"], wall_ms=best["wall_ms"],
               gcups_fill=best["cell_states"] / best["fill_ms"] / 1e6, gcups_total=best["cell_states"] / best["wall_ms"] / 1e6,
               kernel_kind=best["kernel_kind"], warps=best["warps_per_cta"], waves=best["waves"], code_GB=best["code_bytes"] / 1e9)
    print(json.dumps(out), flush=True)
    return out


def main():
    which = sys.argv[1:] or ["1", "2", "3", "4", "5", "na"]
    prot = workloads.PROTEIN_PARAMS
    if "1" in which:
        seqs = ["RAKLPLKEKKLTATANYHPGIRYIMTGYSAKYIYSSTYARFR", "KAKLPLKEKKLTRTANYHPGIRYIMTGYSAKRIYSSTYAYFR"]
        structs = ["CHHHHHHHHHHHHHCCCCTCEEEEEEECCTCEEEEEEEECCC", "HHHHHHHHHHHHCCCCCCTCEEEEEEECCCCCEEEEEEEECC"]
        al = BatchAligner(max_shift=1, **prot)
        res, cls, off = al.encode(seqs, structs)
        run("cfg1 README toy x1", al, res, cls, off, np.array([0], np.int32), np.array([1], np.int32), True)
        run("cfg1 README toy x100000 copies", al, res, cls, off, np.zeros(100000, np.int32), np.ones(100000, np.int32), True)
    if "2" in which:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "dnapol1.json")))
        al = BatchAligner(**g["params"])
        res, cls, off = al.encode([g["seqA"], g["seqB"]], [g["strA"], g["strB"]])
        run("cfg2 DNAPol1 928x933 s=1", al, res, cls, off, np.array([0], np.int32), np.array([1], np.int32), True)
    if "3" in which:
        al = BatchAligner(max_shift=2, **prot)
        res, cls, off, pa, pb = workloads.protein_pairs(12500, seed=3)
        run("cfg3 12.5k protein pairs 200-500 s=2 trace", al, res, cls, off, pa, pb, True)
        run("cfg3 12.5k protein pairs 200-500 s=2 score-only", al, res, cls, off, pa, pb, False)
    if "3s" in which:  # small slice (profiler captures)
        al = BatchAligner(max_shift=2, **prot)
        res, cls, off, pa, pb = workloads.protein_pairs(3552, seed=3)
        run("cfg3 3552 protein pairs 200-500 s=2 trace", al, res, cls, off, pa, pb, True, reps=1)
    if "4s" in which:
        al = BatchAligner(max_shift=2, **workloads.RNA_PARAMS)
        res, cls, off, pa, pb = workloads.rna_pairs(30000, seed=4)
        run("cfg4 30k RNA pairs len 120 s=2 score-only", al, res, cls, off, pa, pb, False, reps=1)
    if "4" in which:
        al = BatchAligner(max_shift=2, **workloads.RNA_PARAMS)
        res, cls, off, pa, pb = workloads.rna_pairs(125000, seed=4)
        run("cfg4 125k RNA pairs len 120 s=2 score-only", al, res, cls, off, pa, pb, False)
        run("cfg4 125k RNA pairs len 120 s=2 trace", al, res, cls, off, pa, pb, True)
    if "na12k" in which:  # the non-affine model on the config-3 slice (12 500 pairs: no tail effect), optionally one CTA width
        na = dict(prot, gap_opening_cost=0, gap_cost=-200, shift_cost=-250)
        al = BatchAligner(max_shift=2, **na)
        if os.environ.get("BA_WARPS"):
            al.set_option("warps_per_cta", int(os.environ["BA_WARPS"]))
        res, cls, off, pa, pb = workloads.protein_pairs(12500, seed=3)
        run("non-affine: 12500 protein pairs 200-500 s=2 trace" + (" G=" + os.environ["BA_WARPS"] if os.environ.get("BA_WARPS") else ""),
            al, res, cls, off, pa, pb, True)
    if "na" in which:
        na = dict(prot, gap_opening_cost=0, gap_cost=-200, shift_cost=-250)
        al = BatchAligner(max_shift=2, **na)
        res, cls, off, pa, pb = workloads.protein_pairs(3000, seed=3)
        run("non-affine: 3000 protein pairs 200-500 s=2 trace", al, res, cls, off, pa, pb, True)
        run("non-affine: 3000 protein pairs 200-500 s=2 score-only", al, res, cls, off, pa, pb, False)
        al.set_option("kernel", 0)
        res, cls, off, pa, pb = workloads.protein_pairs(64, seed=3)
        run("non-affine, general level kernel: 64 pairs trace", al, res, cls, off, pa, pb, True, reps=1)
        al.set_option("kernel", -1)
    if "5" in which:
        al = BatchAligner(max_shift=3, **prot)
        res, cls, off, pa, pb = workloads.protein_pairs(1, lo=8192, hi=8192, seed=5)
        run("cfg5 8192x8192 s=3 trace", al, res, cls, off, pa, pb, True, reps=2)
        run("cfg5 8192x8192 s=3 score-only", al, res, cls, off, pa, pb, False, reps=2)


if __name__ == "__main__":
    main()
