"""Pins the CPU oracle (oracle/bialign_oracle.c) to the reference: README known answers and the
golden vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import oracle


def test_readme_known_answers(golden_cases):
    # README.md:90-104 and README.md:128-152 of the reference
    rna, rna_na, p0, p1, p2 = golden_cases[:5]
    assert rna["score"] == 6800 and rna["trace"] == "faa8fffffffffff2ffff"
    assert rna_na["score"] == 6300
    assert (p0["score"], p1["score"], p2["score"]) == (48300, 48500, 48500)
    assert p1["trace"] == "2ffffffffffffdffffffffffffffffffdffffffff2ff"


@pytest.mark.parametrize("mode", ["literal", "codes"])
def test_oracle_matches_reference_goldens(golden_cases, mode):
    bad = []
    for idx, c in enumerate(golden_cases):
        r = oracle.run(c["seqA"], c["seqB"], c["strA"], c["strB"], c["params"], mode=mode)
        if r["score"] != c["score"] or r["trace"] != c["trace"] or (not r["complete"]) != c["warned"]:
            bad.append((idx, c["params"], c["score"], r["score"], c["trace"], r["trace"]))
    assert not bad, bad[:3]


def test_codes_equal_literal_on_random_mid_size():
    rng = np.random.default_rng(7)
    aa = "ARNDCQEGHILKMFPSTWYV"
    base = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                shift_cost=-150)
    for it in range(12):
        n, m, s = int(rng.integers(20, 60)), int(rng.integers(20, 60)), int(rng.integers(0, 4))
        a = "".join(aa[i] for i in rng.integers(0, 20, n))
        b = "".join(aa[i] for i in rng.integers(0, 20, m))
        sa = "".join("HEC"[i] for i in rng.integers(0, 3, n))
        sb = "".join("HEC"[i] for i in rng.integers(0, 3, m))
        var = [{}, {"shift_cost": 0}, {"structure_weight": 0}, {"gap_cost": 0}][it % 4]
        p = dict(base, max_shift=s, **var)
        lit = oracle.run(a, b, sa, sb, p, mode="literal")
        cod = oracle.run(a, b, sa, sb, p, mode="codes")
        assert lit["score"] == cod["score"] and lit["trace"] == cod["trace"]
        v, end = oracle.eval_trace(a, b, sa, sb, p, lit["trace"])
        assert v == lit["score"] and end == [n, m, n, m]
