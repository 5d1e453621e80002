"""GPU parity: the CUDA path (through the C ABI) against the reference goldens and the CPU oracle."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from bialign_b200 import _capi

    return _capi.get_engine()


def _aligner(params):
    from bialign_b200.batch import BatchAligner

    return BatchAligner(**params)


def test_readme_protein_toy_through_dropin():
    from bialign_b200 import bialignment as ba

    args = {'type': 'Protein', 'gap_cost': -50, 'gap_opening_cost': -150, 'shift_cost': -150,
            'structure_weight': 800, 'max_shift': 1, 'simmatrix': 'BLOSUM62', 'nameA': 'A', 'nameB': 'B'}
    b = ba.BiAligner("RAKLPLKEKKLTATANYHPGIRYIMTGYSAKYIYSSTYARFR", "KAKLPLKEKKLTRTANYHPGIRYIMTGYSAKRIYSSTYAYFR",
                     "CHHHHHHHHHHHHHCCCCTCEEEEEEECCTCEEEEEEEECCC", "HHHHHHHHHHHHCCCCCCTCEEEEEEECCCCCEEEEEEEECC", **args)
    score = b.optimize()
    assert isinstance(score, np.int64) and score == 48500  # README.md:134
    tr = b.traceback()
    assert "".join("%x" % (8 * x[0] + 4 * x[1] + 2 * x[2] + x[3]) for x in tr) == \
        "2ffffffffffffdffffffffffffffffffdffffffff2ff"


def test_goldens_affine(golden_cases):
    """Every affine golden case of the unmodified reference: score and trace bit-exact."""
    from bialign_b200.batch import trace_hex

    groups = {}
    for idx, c in enumerate(golden_cases):
        if c["params"]["gap_opening_cost"] == 0:
            continue
        key = tuple(sorted((k, v) for k, v in c["params"].items()))
        groups.setdefault(key, []).append(idx)
    checked = 0
    for key, idxs in groups.items():
        params = dict(key)
        al = _aligner(params)
        seqs, structs, pairs = [], [], []
        for q, idx in enumerate(idxs):
            c = golden_cases[idx]
            seqs += [c["seqA"], c["seqB"]]
            structs += [c["strA"], c["strB"]]
            pairs.append((2 * q, 2 * q + 1))
        scores, cols, offsets, complete = al.align(seqs, structs, pairs, want_trace=True)
        for q, idx in enumerate(idxs):
            c = golden_cases[idx]
            assert int(scores[q]) == c["score"], (idx, params)
            assert trace_hex(cols, offsets, q) == c["trace"], (idx, params)
            assert bool(complete[q]) != c["warned"]
            checked += 1
        # score-only run must agree too
        s2 = al.align(seqs, structs, pairs, want_trace=False)
        assert (s2 == scores).all()
    assert checked > 100


def _random_protein_batch(rng, npairs, lo, hi):
    aa = "ARNDCQEGHILKMFPSTWYV"
    seqs, structs, pairs = [], [], []
    for q in range(npairs):
        for _ in range(2):
            L = int(rng.integers(lo, hi + 1))
            seqs.append("".join(aa[i] for i in rng.integers(0, 20, L)))
            st = ""
            while len(st) < L:
                st += "HEC"[rng.integers(0, 3)] * int(rng.integers(3, 13))
            structs.append(st[:L])
        pairs.append((2 * q, 2 * q + 1))
    return seqs, structs, pairs


@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_random_batch_vs_oracle(s):
    """Ragged random protein batch vs the CPU oracle: scores, traces and the full code table."""
    from bialign_b200.batch import trace_hex

    rng = np.random.default_rng(100 + s)
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=s)
    seqs, structs, pairs = _random_protein_batch(rng, 24, 1, 70)
    al = _aligner(params)
    scores, cols, offsets, complete = al.align(seqs, structs, pairs, want_trace=True)
    for q, (ia, ib) in enumerate(pairs):
        r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes", want_codes=(q < 6))
        assert int(scores[q]) == r["score"], q
        assert trace_hex(cols, offsets, q) == r["trace"], q
        assert bool(complete[q]) == r["complete"]
        v, end = oracle.eval_trace(seqs[ia], seqs[ib], structs[ia], structs[ib], params, trace_hex(cols, offsets, q))
        assert v == r["score"] and end == [len(seqs[ia]), len(seqs[ib])] * 2
        assert (al.engine.debug_end_values(q) == r["end_values"]).all()
