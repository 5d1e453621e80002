#!/usr/bin/env python3
"""BASELINE config 2 fixture: the DNA polymerase I pair of the reference's Examples/ (CFSSP reports), parsed with
the Query/Struc rule of nonpyx:61-82.  Expected values come from the UNMODIFIED reference (SCORE 761500,
1022 columns, sha256(repr(trace)) = 9f598c58...; an 18-minute Cython run) and are re-derived here with the pinned
CPU oracle in seconds; the script asserts that both agree.  Build container only."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402
from bialign_b200.presentation import read_molecule_from_file  # noqa: E402

a, sa = read_molecule_from_file("/root/reference/Examples/DNAPolymerase1_Escherichia.cfssp", "Protein")
b, sb = read_molecule_from_file("/root/reference/Examples/DNAPolymerase1_Xanthomonas.cfssp", "Protein")
params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
              shift_cost=-150, max_shift=1)  # README.md:159-162
r = oracle.run(a, b, sa, sb, params, mode="codes")
trace = [[int(c, 16) >> 3 & 1, int(c, 16) >> 2 & 1, int(c, 16) >> 1 & 1, int(c, 16) & 1] for c in r["trace"]]
h = hashlib.sha256(repr(trace).encode()).hexdigest()
print(len(a), len(b), r["score"], len(trace), h)
assert r["score"] == 761500 and len(trace) == 1022
assert h == "9f598c582c7355fe5b016a9323df4c837b9fe3f5be7167f279d2bc2a0b8efdeb", "oracle trace differs from the reference run recorded in SURVEY 8c"
json.dump(dict(seqA=a, strA=sa, seqB=b, strB=sb, params=params, score=r["score"], trace=r["trace"],
               trace_repr_sha256=h), open(os.path.join(HERE, "dnapol1.json"), "w"), indent=0)
