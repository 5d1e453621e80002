#!/usr/bin/env python3
"""Generate golden vectors from the UNMODIFIED reference (compiled by oracle/build_ref.sh).

Run in the build container only (needs oracle/_ref):  python tests/golden/make_golden.py
Writes tests/golden/reference_cases.json: inputs, parameters, optimize() score and the traceback()
columns (hex, 8*x0+4*x1+2*x2+x3) for every case.  Deterministic (seeded).
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle", "_ref"))
import bialignment  # noqa: E402  (the reference)

AA = "ARNDCQEGHILKMFPSTWYV"


def rand_protein(rng, L):
    seq = "".join(AA[i] for i in rng.integers(0, 20, L))
    st = ""
    while len(st) < L:
        st += "HEC"[rng.integers(0, 3)] * int(rng.integers(1, 5))
    return seq, st[:L]


def rand_rna(rng, L):
    seq = "".join("ACGU"[i] for i in rng.integers(0, 4, L))
    st, stack = [], []
    for i in range(L):
        r = rng.random()
        if r < 0.3:
            stack.append(i)
            st.append("(")
        elif r < 0.6 and stack and i - stack[-1] >= 1:
            stack.pop()
            st.append(")")
        else:
            st.append(".")
    for i in stack:
        if rng.random() < 0.7:  # leave some '(' unclosed on purpose (treated as unpaired, pyx:378-392)
            st[i] = "."
    return seq, "".join(st)


def run_ref(seqA, seqB, strA, strB, params):
    p = dict(params, nameA="A", nameB="B")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        b = bialignment.BiAligner(seqA, seqB, strA, strB, **p)
        score = int(b.optimize())
        tr = b.traceback()
        full = [[name, row] for name, row in b.decode_trace_full(tr)]
        ev = list(b.eval_trace(tr))
    return dict(seqA=seqA, seqB=seqB, strA=strA, strB=strB, params=params, score=score,
                trace="".join("%x" % (8 * x[0] + 4 * x[1] + 2 * x[2] + x[3]) for x in tr),
                warned="incomplete traceback" in buf.getvalue(), full=full, eval_tail=ev[-2:])


PROT = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
            shift_cost=-150, sequence_match_similarity=100, sequence_mismatch_similarity=0)
RNA = dict(type="RNA", simmatrix=None, structure_weight=400, gap_opening_cost=-200, gap_cost=-50, shift_cost=-150,
           sequence_match_similarity=100, sequence_mismatch_similarity=0)
# tie storms and odd signs
VARIANTS = [{}, {"shift_cost": 0}, {"gap_cost": 0}, {"structure_weight": 0}, {"shift_cost": 0, "gap_cost": 0},
            {"gap_opening_cost": -1}, {"gap_opening_cost": 70}, {"shift_cost": 40, "gap_cost": -300},
            {"gap_opening_cost": 0}, {"gap_opening_cost": 0, "shift_cost": 0}, {"gap_opening_cost": 0, "gap_cost": -200, "shift_cost": -250}]


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    # README known answers (README.md:82-104, 118-152)
    cases.append(run_ref("GCGGGGGAUAUCCCCAUCG", "GGGGAUAUCCCCAUCG", "...(((.....))).....", ".(((.....)))....",
                         dict(RNA, max_shift=1)))
    cases.append(run_ref("GCGGGGGAUAUCCCCAUCG", "GGGGAUAUCCCCAUCG", "...(((.....))).....", ".(((.....)))....",
                         dict(RNA, max_shift=2, gap_opening_cost=0, gap_cost=-200, shift_cost=-250)))
    pa = ("RAKLPLKEKKLTATANYHPGIRYIMTGYSAKYIYSSTYARFR", "KAKLPLKEKKLTRTANYHPGIRYIMTGYSAKRIYSSTYAYFR",
          "CHHHHHHHHHHHHHCCCCTCEEEEEEECCTCEEEEEEEECCC", "HHHHHHHHHHHHCCCCCCTCEEEEEEECCCCCEEEEEEEECC")
    for s in (0, 1, 2):
        cases.append(run_ref(*pa, dict(PROT, max_shift=s)))
    # empty sequences are not golden cases: the reference raises IndexError (seq[-1] on an empty string,
    # pyx:407) as soon as mu1 is evaluated -- in optimize() if exactly one is empty, in traceback() if both
    for it in range(170):
        is_rna = it % 2 == 1
        base = RNA if is_rna else PROT
        var = VARIANTS[it % len(VARIANTS)] if it >= 20 else {}
        s = int(rng.integers(0, 4))
        hi = 13 if s <= 1 else (10 if s == 2 else 8)
        n, m = int(rng.integers(1, hi + 1)), int(rng.integers(1, hi + 1))
        gen = rand_rna if is_rna else rand_protein
        a, sa = gen(rng, n)
        if rng.random() < 0.5 and n and m:  # related pair: mutate A into B
            b = list(a[:m].ljust(m, a[0]))
            sb = list(sa[:m].ljust(m, sa[0] if not is_rna else "."))
            for q in range(m):
                if rng.random() < 0.2:
                    b[q] = gen(rng, 1)[0]
            b, sb = "".join(b), "".join(sb)
            if is_rna:
                sb = rand_rna(rng, m)[1] if rng.random() < 0.5 else sb.replace(")", ".")
        else:
            b, sb = gen(rng, m)
        cases.append(run_ref(a, b, sa, sb, dict(base, max_shift=s, **var)))
        print(it, n, m, s, cases[-1]["score"], cases[-1]["trace"], file=sys.stderr)
    with open(os.path.join(HERE, "reference_cases.json"), "w") as fh:
        json.dump(cases, fh, indent=0)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
