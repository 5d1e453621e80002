#!/usr/bin/env python3
"""Sweep warps_per_cta for a config shape (tuning aid for the engine's per-batch CTA-width model)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bialign_b200 import workloads
from bialign_b200.batch import BatchAligner
import bench_configs as bc

which = sys.argv[1]
for G in range(2, 9):
    if which == "1":
        seqs = ["RAKLPLKEKKLTATANYHPGIRYIMTGYSAKYIYSSTYARFR", "KAKLPLKEKKLTRTANYHPGIRYIMTGYSAKRIYSSTYAYFR"]
        structs = ["CHHHHHHHHHHHHHCCCCTCEEEEEEECCTCEEEEEEEECCC", "HHHHHHHHHHHHCCCCCCTCEEEEEEECCCCCEEEEEEEECC"]
        al = BatchAligner(max_shift=1, **workloads.PROTEIN_PARAMS)
        res, cls, off = al.encode(seqs, structs)
        args = (res, cls, off, np.zeros(100000, np.int32), np.ones(100000, np.int32), True)
    elif which == "4":
        al = BatchAligner(max_shift=2, **workloads.RNA_PARAMS)
        args = workloads.rna_pairs(60000, seed=4) + (False,)
    elif which == "3":
        al = BatchAligner(max_shift=2, **workloads.PROTEIN_PARAMS)
        args = workloads.protein_pairs(3000, seed=3) + (True,)
    elif which == "2":
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "dnapol1.json")))
        al = BatchAligner(**g["params"])
        res, cls, off = al.encode([g["seqA"], g["seqB"]], [g["strA"], g["strB"]])
        args = (res, cls, off, np.array([0], np.int32), np.array([1], np.int32), True)
    else:
        al = BatchAligner(max_shift=3, **workloads.PROTEIN_PARAMS)
        args = workloads.protein_pairs(1, lo=8192, hi=8192, seed=5) + (True,)
    al.set_option("warps_per_cta", G)
    try:
        r = bc.run(f"cfg{which} G={G}", al, *args, reps=2)
    except Exception as ex:
        print("G", G, "failed", ex)
    al.set_option("warps_per_cta", 0)
