#!/usr/bin/env python3
"""Long-pair flavour with and without the I/O warp (engine option io_warp) on the shapes that use it (tuning aid)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bialign_b200 import workloads
from bialign_b200.batch import BatchAligner
import bench_configs as bc

prot = workloads.PROTEIN_PARAMS
which = sys.argv[1:] or ["2", "5", "lb", "wr"]
for iow in (0, 1):
    if "2" in which:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "dnapol1.json")))
        al = BatchAligner(**g["params"]); al.set_option("io_warp", iow)
        res, cls, off = al.encode([g["seqA"], g["seqB"]], [g["strA"], g["strB"]])
        bc.run(f"io_warp={iow} cfg2 DNAPol1", al, res, cls, off, np.array([0], np.int32), np.array([1], np.int32), True, reps=5)
    if "5" in which:
        al = BatchAligner(max_shift=3, **prot); al.set_option("io_warp", iow)
        bc.run(f"io_warp={iow} cfg5 8192^2 s=3", al, *workloads.protein_pairs(1, lo=8192, hi=8192, seed=5), True, reps=2)
    if "lb" in which:
        al = BatchAligner(max_shift=2, **prot); al.set_option("io_warp", iow)
        bc.run(f"io_warp={iow} 400 pairs 1900-2100 s=2", al, *workloads.protein_pairs(400, lo=1900, hi=2100, seed=6), True, reps=2)
    if "wr" in which:
        al = BatchAligner(max_shift=2, **dict(prot, structure_weight=333, gap_opening_cost=-157, gap_cost=-49, shift_cost=-151)); al.set_option("io_warp", iow)
        bc.run(f"io_warp={iow} wide range 400 pairs 1900-2100 s=2", al, *workloads.protein_pairs(400, lo=1900, hi=2100, seed=6), True, reps=1)
