"""GPU checks on the BASELINE.json configurations themselves (through the C ABI).
Full-size inputs are checked through size-independent properties -- every emitted trace re-scores
(column by column, independent of any DP table, pyx:745-800) to the reported score and ends at (n,m,n,m);
score-only and score+trace runs agree -- plus oracle comparisons on a bounded subsample."""
import hashlib
import json
import os

import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _hex(cols, offsets, p):
    return "".join("%x" % c for c in cols[offsets[p]:offsets[p + 1]])


def test_config1_readme_protein_toy_cli_equivalent():
    from bialign_b200.batch import BatchAligner

    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=1)
    al = BatchAligner(**params)
    seqs = ["RAKLPLKEKKLTATANYHPGIRYIMTGYSAKYIYSSTYARFR", "KAKLPLKEKKLTRTANYHPGIRYIMTGYSAKRIYSSTYAYFR"]
    structs = ["CHHHHHHHHHHHHHCCCCTCEEEEEEECCTCEEEEEEEECCC", "HHHHHHHHHHHHCCCCCCTCEEEEEEECCCCCEEEEEEEECC"]
    scores, cols, offsets, complete = al.align(seqs * 64, structs * 64, [(2 * q, 2 * q + 1) for q in range(64)], want_trace=True)
    assert (scores == 48500).all() and al.engine.stats()["cell_states"] == 64 * 145161
    assert all(_hex(cols, offsets, q) == "2ffffffffffffdffffffffffffffffffdffffffff2ff" for q in range(64))


def test_config2_dnapol1_pair_full_traceback():
    """928 x 933 aa, max_shift 1: SCORE 761500 and the exact 1022-column trace of the reference run."""
    from bialign_b200.batch import BatchAligner

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "dnapol1.json")))
    al = BatchAligner(**g["params"])
    for long_mode in (1, 0):
        al.set_option("long", long_mode)
        try:
            scores, cols, offsets, complete = al.align([g["seqA"], g["seqB"]], [g["strA"], g["strB"]], [(0, 1)], want_trace=True)
        finally:
            al.set_option("long", -1)
        assert int(scores[0]) == 761500 == g["score"]
        tr = _hex(cols, offsets, 0)
        assert tr == g["trace"] and bool(complete[0])
        trace = [[int(c, 16) >> 3 & 1, int(c, 16) >> 2 & 1, int(c, 16) >> 1 & 1, int(c, 16) & 1] for c in tr]
        assert hashlib.sha256(repr(trace).encode()).hexdigest() == g["trace_repr_sha256"]
        assert al.engine.stats()["cell_states"] == 70182000


def _rescore_all(res, cls, off, pa, pb, params, scores, cols, offsets, decode, step=1):
    bad = 0
    for p in range(0, len(pa), step):
        a, sa = decode(res, cls, off, int(pa[p]))
        b, sb = decode(res, cls, off, int(pb[p]))
        v, end = oracle.eval_trace(a, b, sa, sb, params, _hex(cols, offsets, p))
        bad += (v != int(scores[p])) or (end != [len(a), len(b), len(a), len(b)])
    return bad


def test_config3_slice_properties_and_subsample():
    """2 000 pairs of the config-3 generator (seed 3): all traces re-score to their scores; 6 pairs vs the oracle."""
    from bialign_b200 import workloads
    from bialign_b200.batch import BatchAligner

    params = dict(workloads.PROTEIN_PARAMS, max_shift=2)
    res, cls, off, pa, pb = workloads.protein_pairs(2000, seed=3)
    al = BatchAligner(**params)
    scores, cols, offsets, complete = al.align_encoded(res, cls, off, pa, pb, want_trace=True)
    assert complete.all() and al.engine.stats()["kernel_kind"] == 1
    assert (al.align_encoded(res, cls, off, pa, pb, want_trace=False) == scores).all()
    assert al.engine.stats()["kernel_kind"] == 1  # 200-500 aa with BLOSUM62: does not fit the 16-bit pair mode
    assert _rescore_all(res, cls, off, pa, pb, params, scores, cols, offsets, workloads.decode_protein, step=5) == 0
    for p in (0, 1, 777, 1234, 1998, 1999):
        a, sa = workloads.decode_protein(res, cls, off, int(pa[p]))
        b, sb = workloads.decode_protein(res, cls, off, int(pb[p]))
        r = oracle.run(a, b, sa, sb, params, mode="codes")
        assert int(scores[p]) == r["score"] and _hex(cols, offsets, p) == r["trace"], p


def test_config4_rna_slice_score_only():
    """20 000 pairs of the config-4 generator (seed 4, length 120, supplied dot-bracket), score only."""
    from bialign_b200 import workloads
    from bialign_b200.batch import BatchAligner

    params = dict(workloads.RNA_PARAMS, max_shift=2)
    res, cls, off, pa, pb = workloads.rna_pairs(20000, seed=4)
    al = BatchAligner(**params)
    al.table = __import__("bialign_b200.encoding", fromlist=["x"]).match_table(100, 0, nsym=4)
    scores = al.align_encoded(res, cls, off, pa, pb, want_trace=False)
    assert al.engine.stats()["cell_states"] == 20000 * 3229209  # 3 229 209 cell-states per pair
    assert al.engine.stats()["kernel_kind"] == 5  # config 4 runs two pairs per lane in packed 16-bit halves
    al.set_option("p16", 0)
    try:
        assert (al.align_encoded(res, cls, off, pa, pb, want_trace=False) == scores).all()  # 32-bit kernel agrees
    finally:
        al.set_option("p16", -1)
    for p in list(range(0, 20000, 667)) + [19999]:
        a, sa = workloads.decode_rna(res, cls, off, int(pa[p]))
        b, sb = workloads.decode_rna(res, cls, off, int(pb[p]))
        assert int(scores[p]) == oracle.run(a, b, sa, sb, params, mode="codes")["score"], p
    # traces of a sub-slice re-score to the scores
    s2, cols, offsets, complete = al.align_encoded(res, cls, off, pa[:500], pb[:500], want_trace=True)
    assert (s2 == scores[:500]).all() and complete.all()
    assert _rescore_all(res, cls, off, pa[:500], pb[:500], params, s2, cols, offsets, workloads.decode_rna, step=7) == 0


@pytest.mark.parametrize("name", ["cfg5_small", "cfg5"])
def test_config5_long_pair(name):
    """One long protein pair, max_shift 3, multi-CTA fill with the traceback codes in HBM (8192 x 8192: 26 GB).
    Golden score / trace hash from the CPU oracle (tests/golden/make_long_golden.py)."""
    from bialign_b200 import workloads
    from bialign_b200.batch import BatchAligner

    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "long_pairs.json")))
    shapes = {"cfg5_small": dict(length=1024, max_shift=3, seed=5), "cfg5": dict(length=8192, max_shift=3, seed=5)}
    g = gold.get(name, shapes[name])  # without a golden entry only the size-independent properties are checked
    params = dict(workloads.PROTEIN_PARAMS, max_shift=g["max_shift"])
    res, cls, off, pa, pb = workloads.protein_pairs(1, lo=g["length"], hi=g["length"], seed=g["seed"])
    al = BatchAligner(**params)
    scores, cols, offsets, complete = al.align_encoded(res, cls, off, pa, pb, want_trace=True)
    st = al.engine.stats()
    assert st["kernel_kind"] in (3, 4)
    tr = _hex(cols, offsets, 0)
    a, sa = workloads.decode_protein(res, cls, off, 0)
    b, sb = workloads.decode_protein(res, cls, off, 1)
    v, end = oracle.eval_trace(a, b, sa, sb, params, tr)
    assert v == int(scores[0]) and end == [g["length"]] * 4 and bool(complete[0])
    if "score" in g:
        assert int(scores[0]) == g["score"] and len(tr) == g["trace_len"]
        assert hashlib.sha256(tr.encode()).hexdigest() == g["trace_sha256"]
    print(name, "fill_ms", st["fill_ms"], "traceback_ms", st["traceback_ms"], "GCUPS", st["cell_states"] / st["fill_ms"] / 1e6)
