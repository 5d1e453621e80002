#!/usr/bin/env python3
"""Golden vectors for the PROBABILISTIC structure similarity (pyx:345-353, 414-423) from the UNMODIFIED reference.

ViennaRNA is not installed here, so the reference is given tests/fake_rna/RNA.py (a deterministic stand-in with the four calls
the reference makes) and run with strA = strB = None, or with one supplied and one predicted structure.  Build container only
(needs oracle/_ref):   python tests/golden/make_rna_prob_golden.py   ->  tests/golden/rna_prob_cases.json"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "fake_rna"))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle", "_ref"))
import bialignment  # noqa: E402  (the reference)

RNA_PARAMS = dict(type="RNA", simmatrix=None, structure_weight=400, gap_opening_cost=-200, gap_cost=-50, shift_cost=-150,
                  sequence_match_similarity=100, sequence_mismatch_similarity=0)


def run(seqA, seqB, strA, strB, params):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        b = bialignment.BiAligner(seqA, seqB, strA, strB, nameA="A", nameB="B", **params)
        score = int(b.optimize())
        tr = b.traceback()
        full = [[n, r] for n, r in b.decode_trace_full(tr)]
        ev = list(b.eval_trace(tr))
    return dict(seqA=seqA, seqB=seqB, strA=strA, strB=strB, params=params, score=score,
                trace="".join("%x" % (8 * x[0] + 4 * x[1] + 2 * x[2] + x[3]) for x in tr), full=full, eval_tail=ev[-2:],
                warned="incomplete traceback" in buf.getvalue())


def main():
    rng = np.random.default_rng(424242)
    cases = []
    for q in range(14):
        la, lb = int(rng.integers(8, 34)), int(rng.integers(8, 34))
        a = "".join("ACGU"[i] for i in rng.integers(0, 4, la))
        b = "".join("ACGU"[i] for i in rng.integers(0, 4, lb))
        p = dict(RNA_PARAMS, max_shift=int(rng.integers(0, 3)))
        if q % 4 == 1:
            p["gap_opening_cost"] = 0       # non-affine model
            p["gap_cost"], p["shift_cost"] = -200, -250
        if q % 5 == 2:
            p["structure_weight"] = 333     # not a multiple of anything: exercises the float -> int truncation
        strA = None
        strB = None if q % 3 else "." * lb  # mixed: predicted A, supplied (unpaired) B
        cases.append(run(a, b, strA, strB, p))
        print(q, la, lb, p["max_shift"], cases[-1]["score"], file=sys.stderr)
    json.dump(cases, open(os.path.join(HERE, "rna_prob_cases.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
