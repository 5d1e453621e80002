#!/usr/bin/env python3
"""Randomized differential run: random parameters / max_shift / lengths / kernels against the CPU oracle's literal
restatement.    python scripts/fuzz_parity.py [seed] [rounds]
A 100-round slice runs inside the GPU test suite (tests/test_gpu_parity.py::test_fuzz_slice_vs_literal_oracle)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

AA = "ARNDCQEGHILKMFPSTWYV"


def fuzz(seed=1, rounds=40, verbose=True, max_len=160):
    """Returns (mismatches, pairs checked).  Every round: random scoring parameters (incl. zero and positive costs),
    RNA or protein, max_shift 0..4, 1-8 pairs of lengths 0..max_len, a random kernel / flavour / CTA width / long-pair
    mode / rebased trace run; scores, traces and completeness flags must equal the oracle's literal int64 restatement."""
    import oracle
    from bialign_b200.batch import BatchAligner, trace_hex

    rng = np.random.default_rng(seed)
    bad = checked = 0
    for rd in range(rounds):
        s = int(rng.integers(0, 5))
        rna = rng.random() < 0.4
        step = int(rng.choice([1, 5, 10, 50]))
        params = dict(type="RNA" if rna else "Protein", simmatrix=None if rna else "BLOSUM62",
                      structure_weight=int(rng.integers(0, 20)) * step, gap_opening_cost=-int(rng.integers(0, 8)) * step,
                      gap_cost=-int(rng.integers(0, 8)) * step,
                      shift_cost=-int(rng.integers(0, 8)) * step + (step if rng.random() < 0.1 else 0), max_shift=s,
                      sequence_match_similarity=int(rng.integers(0, 5)) * step,
                      sequence_mismatch_similarity=-int(rng.integers(0, 3)) * step)
        if rng.random() < 0.1:
            params["gap_opening_cost"] = int(rng.integers(1, 4)) * step  # positive opening
        hi = int(rng.choice([12, 40, 90, max_len]))
        npairs = int(rng.integers(1, 9))
        seqs, structs, pairs = [], [], []
        for q in range(npairs):
            for _ in range(2):
                L = int(rng.integers(0 if rng.random() < 0.05 else 1, hi + 1))
                if rna:
                    seqs.append("".join("ACGU"[i] for i in rng.integers(0, 4, L)))
                    st, stack = [], []
                    for i in range(L):
                        u = rng.random()
                        if u < 0.3:
                            stack.append(i)
                            st.append("(")
                        elif u < 0.6 and stack:
                            stack.pop()
                            st.append(")")
                        else:
                            st.append(".")
                    structs.append("".join(st))
                else:
                    seqs.append("".join(AA[i] for i in rng.integers(0, 20, L)))
                    structs.append("".join("HEC"[i] for i in rng.integers(0, 3, L)))
            pairs.append((2 * q, 2 * q + 1))
        kern = int(rng.choice([-1, -1, 0, 1]))
        al = BatchAligner(**params)
        al.set_option("kernel", kern if kern != 1 else -1)
        al.set_option("pad", int(rng.choice([-1, 0, 1])))
        al.set_option("long", int(rng.choice([-1, -1, 1])))
        al.set_option("warps_per_cta", int(rng.choice([0, 0, 2, 4, 6])))
        al.set_option("io_warp", int(rng.choice([-1, 0, 1])))        # long-pair mode: with / without the I/O warp
        al.set_option("col_chunks", int(rng.choice([0, 0, 2, 3])))   # ... and (single pairs) column-chunked tiles
        if rng.random() < 0.25:  # rebased trace run (forced), sometimes with a window so small that pairs fall back
            al.set_option("rebase", 1)
            al.set_option("rebase_window", int(rng.choice([0, 0, 30, 300])))
        try:
            scores, cols, offsets, complete = al.align(seqs, structs, pairs, want_trace=True)
            kind = al.engine.stats()["kernel_kind"]
            s2 = al.align(seqs, structs, pairs, want_trace=False)
            kind2 = al.engine.stats()["kernel_kind"]
        except Exception as ex:
            if "range" in str(ex) or "requested" in str(ex):
                continue  # forced flavour not applicable to these parameters
            raise
        for q, (ia, ib) in enumerate(pairs):
            r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="literal")
            ok = int(scores[q]) == r["score"] and trace_hex(cols, offsets, q) == r["trace"] and int(s2[q]) == r["score"]
            if params["gap_opening_cost"] != 0:
                ok = ok and (bool(complete[q]) == r["complete"])
            checked += 1
            if not ok:
                bad += 1
                print("MISMATCH round", rd, "pair", q, params, "kinds", kind, kind2, len(seqs[ia]), len(seqs[ib]),
                      int(scores[q]), r["score"], int(s2[q]), trace_hex(cols, offsets, q) == r["trace"], flush=True)
        if verbose:
            print(f"round {rd}: s={s} {'RNA' if rna else 'prot'} npairs={npairs} hi={hi} kinds={kind}/{kind2} ok", flush=True)
    return bad, checked


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    t0 = time.time()
    bad, checked = fuzz(seed, rounds)
    print("DONE rounds", rounds, "pairs", checked, "mismatches", bad, "time %.0f s" % (time.time() - t0))
