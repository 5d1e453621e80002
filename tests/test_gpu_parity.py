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


def test_goldens_all_models(golden_cases):
    """Every golden case of the unmodified reference (affine and non-affine): score and trace bit-exact."""
    from bialign_b200.batch import trace_hex

    groups = {}
    for idx, c in enumerate(golden_cases):
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


def _decode_codes(words, kind):
    """Device code words -> array [cells, 9] of case ids (15 = none)."""
    w = words.astype(np.uint64)
    out = np.zeros((w.size, 9), dtype=np.int64)
    for t in range(9):
        if kind == 0:
            out[:, t] = ((w >> np.uint64(4 * t)) & np.uint64(15)).astype(np.int64)
        else:
            sh = 2 + 5 * t if t < 6 else 32 + 17 + 5 * (t - 6)
            f = ((w >> np.uint64(sh)) & np.uint64(31)).astype(np.int64)
            ids = np.full(f.shape, 15, dtype=np.int64)
            t01, t23 = t // 3, t % 3
            full = (f >= 19) & (f <= 27)
            ids[full] = 27 - f[full]
            rk = 17 - 3 * t01 - f
            h2 = (f >= 9) & (f <= 17) & (rk >= 0) & (rk <= 2)
            ids[h2] = 9 + (2 - rk[h2])
            num = 8 - t23 - f
            h1 = (f <= 8) & (num >= 0) & (num <= 6) & (num % 3 == 0)
            ids[h1] = 12 + (2 - num[h1] // 3)
            out[:, t] = ids
    return out


def _check_batch(al, seqs, structs, pairs, params, table_pairs=0):
    from bialign_b200.batch import trace_hex

    scores, cols, offsets, complete = al.align(seqs, structs, pairs, want_trace=True)
    al.trace_run_stats = al.engine.stats()
    kind = al.trace_run_stats["kernel_kind"]
    for q, (ia, ib) in enumerate(pairs):
        r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes", want_codes=True)
        assert int(scores[q]) == r["score"], (q, len(seqs[ia]), len(seqs[ib]))
        assert trace_hex(cols, offsets, q) == r["trace"], (q, len(seqs[ia]), len(seqs[ib]))
        assert bool(complete[q]) == r["complete"]
        v, end = oracle.eval_trace(seqs[ia], seqs[ib], structs[ia], structs[ib], params, trace_hex(cols, offsets, q))
        assert v == r["score"] and end == [len(seqs[ia]), len(seqs[ib])] * 2
        ev, ov = al.engine.debug_end_values(q).astype(np.int64), r["end_values"].astype(np.int64)
        fin = ov > -(1 << 29)
        assert (ev[fin] == ov[fin]).all() and (ev[~fin] < ov[fin].min()).all()
        if q < table_pairs:
            # every reachable cell-state's winning case, not only those on the optimal path
            try:
                words = al.engine.debug_codes(q, r["codes"].size)
            except Exception:
                continue  # pair not in the last wave
            oc = r["codes"]
            valid = oc != np.uint64(0xFFFFFFFFFFFFFFFF)
            want = _decode_codes(oc[valid], 0)
            got = _decode_codes(words[valid], kind)
            reach = np.stack([((oc[valid] >> np.uint64(36 + t)) & np.uint64(1)).astype(bool) for t in range(9)], axis=1)
            reach &= want != 15
            assert (want[reach] == got[reach]).all(), q
    # score-only run (no tie-break bits, no code stores; 16-bit pair mode when the range fits) must give the same scores
    s2 = al.align(seqs, structs, pairs, want_trace=False)
    assert (s2 == scores).all()
    return kind


def _select(al, kind):
    """kind 0 = generic level kernel, 1 = systolic pad-free, 2 = systolic padded."""
    al.set_option("kernel", 0 if kind == 0 else 1)
    al.set_option("pad", -1 if kind == 0 else kind - 1)
    al.set_option("chain", 0)  # the chained short-pair flavour has its own test


def _unselect(al):
    al.options.clear()  # options belong to the aligner; the next configure() resets the shared engine to automatic


@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("s", [0, 1, 2, 3, 4])
def test_random_batch_vs_oracle(s, kernel):
    """Ragged random protein batch vs the CPU oracle: scores, traces, end values and code tables."""
    rng = np.random.default_rng(100 + s)
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=s)
    seqs, structs, pairs = _random_protein_batch(rng, 24, 1, 70)
    al = _aligner(params)
    _select(al, kernel)
    try:
        assert _check_batch(al, seqs, structs, pairs, params, table_pairs=24) == kernel
    finally:
        _unselect(al)


@pytest.mark.parametrize("pad", [0, 1])
@pytest.mark.parametrize("warps", [1, 2, 4, 8])
def test_systolic_multipass_and_cta_shapes(warps, pad):
    """Pairs longer than one row block (several passes through the boundary stream), all CTA widths."""
    rng = np.random.default_rng(500 + warps)
    for s, tie in ((2, {}), (1, {"shift_cost": 0}), (3, {"structure_weight": 0, "gap_cost": 0})):
        params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150,
                      gap_cost=-50, shift_cost=-150, max_shift=s)
        params.update(tie)
        seqs, structs, pairs = _random_protein_batch(rng, 6, 60, 150)
        al = _aligner(params)
        _select(al, 1 + pad)
        al.set_option("warps_per_cta", warps)
        al.set_option("long", 0)  # (few pairs with many row blocks would otherwise run as long-pair gangs: tested below)
        try:
            assert _check_batch(al, seqs, structs, pairs, params, table_pairs=2) == 1 + pad
        finally:
            _unselect(al)


def test_positive_gap_opening_and_tie_storms():
    """beta > 0 (general open() reduction, padded flavour) and all-zero costs (every case ties)."""
    rng = np.random.default_rng(77)
    for var in ({"gap_opening_cost": 70}, {"shift_cost": 0, "gap_cost": 0, "structure_weight": 0, "gap_opening_cost": -1},
                {"shift_cost": 40, "gap_cost": -300}):
        params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150,
                      gap_cost=-50, shift_cost=-150, max_shift=2)
        params.update(var)
        seqs, structs, pairs = _random_protein_batch(rng, 8, 1, 60)
        al = _aligner(params)
        assert _check_batch(al, seqs, structs, pairs, params, table_pairs=8) in (1, 2)


def test_rna_batch_score_only_vs_oracle():
    """cfg4-shaped: RNA with supplied dot-bracket structures, match/mismatch similarity, score only."""
    rng = np.random.default_rng(4)
    params = dict(type="RNA", simmatrix=None, structure_weight=400, gap_opening_cost=-200, gap_cost=-50,
                  shift_cost=-150, max_shift=2, sequence_match_similarity=100, sequence_mismatch_similarity=0)
    seqs, structs, pairs = [], [], []
    for q in range(16):
        for _ in range(2):
            L = int(rng.integers(30, 121))
            seqs.append("".join("ACGU"[i] for i in rng.integers(0, 4, L)))
            st, stack = [], []
            for i in range(L):
                u = rng.random()
                if u < 0.3:
                    stack.append(i); st.append("(")
                elif u < 0.6 and stack and i - stack[-1] >= 3:
                    stack.pop(); st.append(")")
                else:
                    st.append(".")
            for i in stack:
                st[i] = "."
            structs.append("".join(st))
        pairs.append((2 * q, 2 * q + 1))
    al = _aligner(params)
    scores = al.align(seqs, structs, pairs, want_trace=False)
    for q, (ia, ib) in enumerate(pairs):
        r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes")
        assert int(scores[q]) == r["score"], q


def test_cli_transcripts_match_reference(capsys, monkeypatch):
    """bin/bialign.py stdout == the reference CLI's stdout (tests/golden/cli_outputs.json), all output modes, -v,
    and --fileinput on two CFSSP-format files (tests/golden/small_*.cfssp; the CLI runs from that directory)."""
    import importlib.util
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bialign_cli", os.path.join(root, "bin", "bialign.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    cases = json.load(open(os.path.join(root, "tests", "golden", "cli_outputs.json")))
    assert any("--filein" in c["argv"] for c in cases)
    monkeypatch.chdir(os.path.join(root, "tests", "golden"))
    for c in cases:
        capsys.readouterr()
        try:
            cli.main(c["argv"])
        except SystemExit:
            pass
        out = capsys.readouterr().out
        assert out == c["stdout"], c["argv"]


def test_dropin_nonaffine_and_rna_api():
    from bialign_b200 import bialignment as ba

    b = ba.BiAligner("GCGGGGGAUAUCCCCAUCG", "GGGGAUAUCCCCAUCG", "...(((.....))).....", ".(((.....)))....", type="RNA",
                     simmatrix=None, structure_weight=400, gap_opening_cost=0, gap_cost=-200, shift_cost=-250,
                     max_shift=2, sequence_match_similarity=100, sequence_mismatch_similarity=0, nameA="A", nameB="B")
    assert b.optimize() == 6300
    tr = b.traceback()
    assert all(isinstance(x, tuple) for x in tr)  # non-affine traces are tuples (pyx:526)
    assert "".join("%x" % (8 * x[0] + 4 * x[1] + 2 * x[2] + x[3]) for x in tr) == "aaffffffffffffaffff"
    with pytest.raises(KeyError):
        ba.BiAligner("ACDJ", "ACD", "HHHH", "HHH", type="Protein", simmatrix="BLOSUM62", structure_weight=800,
                     gap_opening_cost=-150, gap_cost=-50, shift_cost=-150, max_shift=1).optimize()


@pytest.mark.parametrize("pad", [0, 1])
@pytest.mark.parametrize("warps", [1, 4])
def test_long_pair_mode_multi_cta(warps, pad):
    """Row blocks of one pair spread over several CTAs (cooperative launch, flag-synchronised boundary streams)."""
    rng = np.random.default_rng(900 + warps + 10 * pad)
    for s in (1, 2, 3):
        params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150,
                      gap_cost=-50, shift_cost=-150, max_shift=s)
        seqs, structs, pairs = _random_protein_batch(rng, 3, 90, 260)
        al = _aligner(params)
        _select(al, 1 + pad)
        al.set_option("warps_per_cta", warps)
        al.set_option("long", 1)
        try:
            assert _check_batch(al, seqs, structs, pairs, params, table_pairs=3) == 3 + pad
        finally:
            _unselect(al)


@pytest.mark.parametrize("warps", [2, 4])
def test_long_pair_io_warp_and_column_chunks(warps):
    """Long-pair mode with the I/O warp (a warp of its own for boundary streams and flags), alone and with the row blocks cut
    into column chunks (tiles dealt in start order, column state handed on through the global column buffer): single pairs
    of several row blocks vs the oracle -- scores, traces, end values and the winning case of every reachable cell-state."""
    rng = np.random.default_rng(4200 + warps)
    for s in (0, 1, 2, 3, 4):
        params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150,
                      gap_cost=-50, shift_cost=-150, max_shift=s)
        for chunks in (1, 2, 3):
            seqs, structs, pairs = _random_protein_batch(rng, 1, 150, 230)
            al = _aligner(params)
            _select(al, 1)
            al.set_option("warps_per_cta", warps)
            al.set_option("long", 1)
            al.set_option("io_warp", 1)
            al.set_option("col_chunks", chunks)
            try:
                assert _check_batch(al, seqs, structs, pairs, params, table_pairs=1) == 3
            finally:
                _unselect(al)
    # a handful of pairs in one launch (gangs) with the I/O warp forced; tie storms through the chunk boundaries
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=0, gap_opening_cost=-1, gap_cost=0, shift_cost=0, max_shift=2)
    seqs, structs, pairs = _random_protein_batch(rng, 3, 100, 180)
    al = _aligner(params)
    _select(al, 1)
    al.set_option("long", 1)
    al.set_option("io_warp", 1)
    try:
        _check_batch(al, seqs, structs, pairs, params, table_pairs=3)
        al.set_option("col_chunks", 2)
        _check_batch(al, seqs[:2], structs[:2], pairs[:1], params, table_pairs=1)
    finally:
        _unselect(al)


def test_long_pair_auto_mode_many_passes():
    """One 1500 x 1400 pair, max_shift 1: auto-selected long mode, dozens of passes over dozens of CTAs."""
    rng = np.random.default_rng(1234)
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=1)
    aa = "ARNDCQEGHILKMFPSTWYV"
    a = "".join(aa[i] for i in rng.integers(0, 20, 1500))
    b = list(a[:1400])
    for q in range(0, 1400, 7):
        b[q] = aa[rng.integers(0, 20)]
    b = "".join(b[:700] + b[650:])[:1400]
    sa = "".join("HEC"[i] for i in rng.integers(0, 3, 1500))
    sb = sa[3:1403]
    al = _aligner(params)
    from bialign_b200.batch import trace_hex
    scores, cols, offsets, complete = al.align([a, b], [sa, sb], [(0, 1)], want_trace=True)
    assert al.engine.stats()["kernel_kind"] in (3, 4)
    r = oracle.run(a, b, sa, sb, params, mode="codes")
    assert int(scores[0]) == r["score"] and trace_hex(cols, offsets, 0) == r["trace"] and bool(complete[0])
    assert (al.align([a, b], [sa, sb], [(0, 1)], want_trace=False) == scores).all()


@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_edge_cases_empty_and_tiny_sequences(kernel):
    """Empty molecules (the batch API defines them by the recurrence; the reference class raises IndexError),
    length-1 molecules, and very unequal lengths."""
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=2)
    seqs = ["", "A", "AC", "ACDEFGHIKLMNPQRSTVWY" * 4, "W", ""]
    structs = ["", "H", "HE", "HHHHHEEEEECCCCCHHHHH" * 4, "C", ""]
    pairs = [(0, 5), (0, 1), (1, 0), (1, 4), (1, 1), (2, 3), (3, 2), (3, 3), (0, 3), (3, 0)]
    al = _aligner(params)
    _select(al, kernel)
    try:
        from bialign_b200.batch import trace_hex

        scores, cols, offsets, complete = al.align(seqs, structs, pairs, want_trace=True)
        for q, (ia, ib) in enumerate(pairs):
            r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes")
            assert int(scores[q]) == r["score"], (q, pairs[q])
            assert trace_hex(cols, offsets, q) == r["trace"], (q, pairs[q])
            assert bool(complete[q]) == r["complete"], (q, pairs[q])
        assert (al.align(seqs, structs, pairs, want_trace=False) == scores).all()
        empty = al.align(seqs, structs, [], want_trace=True)
        assert len(empty[0]) == 0 and empty[2].tolist() == [0]
    finally:
        _unselect(al)


def test_error_codes_range_and_alphabet():
    from bialign_b200 import _capi
    from bialign_b200.batch import BatchAligner

    # a score bound beyond int32: the reference's tables are int64 (pyx:27-35) -> 64-bit level kernel in automatic mode,
    # BA_ERR_SCORE_RANGE only when a fast kernel is forced
    from bialign_b200.batch import trace_hex

    for gap_open in (-150, 0):
        wide = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=1 << 27, gap_opening_cost=gap_open, gap_cost=-50,
                    shift_cost=-(1 << 24), max_shift=1)
        big = BatchAligner(**wide)
        sa, sb, ta, tb = "ACDEFGHIKL" * 6, "ACDFGHIKLM" * 6 + "AC", "HHHEEECCCH" * 6, "HHEEECCCHH" * 6 + "EE"
        scores, cols, offsets, complete = big.align([sa, sb], [ta, tb], [(0, 1)], want_trace=True)
        assert big.engine.stats()["kernel_kind"] == 9
        r = oracle.run(sa, sb, ta, tb, wide, mode="literal")
        assert int(scores[0]) == r["score"] and abs(r["score"]) > (1 << 31) and trace_hex(cols, offsets, 0) == r["trace"]
        assert int(big.align([sa, sb], [ta, tb], [(0, 1)])[0]) == r["score"]
        big.set_option("kernel", 1)
        with pytest.raises(_capi.BialignError) as ei:
            big.align([sa, sb], [ta, tb], [(0, 1)])
        assert ei.value.code == _capi.BA_ERR_SCORE_RANGE
    ok = BatchAligner(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                      shift_cost=-150, max_shift=1)
    res = np.array([0, 1, 200, 3], dtype=np.uint8)  # 200 >= nsym (24): the reference would raise KeyError (pyx:407)
    with pytest.raises(_capi.BialignError) as ei:
        ok.align_encoded(res, np.zeros(4, np.uint8), np.array([0, 2, 4], np.int64), np.array([0], np.int32), np.array([1], np.int32))
    assert ei.value.code == _capi.BA_ERR_ALPHABET
    with pytest.raises(_capi.BialignError) as ei:  # pair index outside the sequence table
        ok.align_encoded(res[:2], np.zeros(2, np.uint8), np.array([0, 1, 2], np.int64), np.array([0], np.int32), np.array([7], np.int32))
    assert ei.value.code == _capi.BA_ERR_INVALID_ARG
    for key, bad in [("kernel", 2), ("pad", -2), ("long", 5), ("p16", 2), ("warps_per_cta", 9), ("code_arena_bytes", -1),
                     ("no_such_option", 0)]:
        with pytest.raises(_capi.BialignError) as ei:
            ok.set_option(key, bad)
        assert ei.value.code == _capi.BA_ERR_INVALID_ARG, key
    with pytest.raises(_capi.BialignError) as ei:  # offsets that are not monotone
        ok.align_encoded(res[:2], np.zeros(2, np.uint8), np.array([0, 2, 1], np.int64), np.array([0], np.int32), np.array([1], np.int32))
    assert ei.value.code == _capi.BA_ERR_INVALID_ARG
    # the engine is still usable after the rejected calls
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=1)
    want = oracle.run("ACDW", "ACW", "HHHC", "HHC", params, mode="codes")["score"]
    assert ok.align(["ACDW", "ACW"], ["HHHC", "HHC"], [(0, 1)], want_trace=False).tolist() == [want]


@pytest.mark.parametrize("s", [0, 1, 2, 3, 4])
def test_score_only_16bit_pair_mode(s):
    """Two pairs per lane in packed 16-bit halves (score-only batches whose range provably fits): ragged
    partners, odd batch size, multi-pass, RNA and protein scoring -- against the oracle and the 32-bit kernel."""
    rng = np.random.default_rng(1600 + s)
    cases = [dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=s),
             dict(type="RNA", simmatrix=None, structure_weight=400, gap_opening_cost=-200, gap_cost=-50, shift_cost=-150,
                  max_shift=s, sequence_match_similarity=100, sequence_mismatch_similarity=0)]
    for params in cases:
        if params["type"] == "Protein":
            seqs, structs, pairs = _random_protein_batch(rng, 13, 1, 90)
        else:
            seqs, structs, pairs = [], [], []
            for q in range(13):
                for _ in range(2):
                    L = int(rng.integers(1, 130))
                    seqs.append("".join("ACGU"[i] for i in rng.integers(0, 4, L)))
                    structs.append("".join(".()"[i] if False else "." for i in range(L)))
                pairs.append((2 * q, 2 * q + 1))
            # simple balanced structures
            structs = [("(" * (len(x) // 3) + "." * (len(x) - 2 * (len(x) // 3)) + ")" * (len(x) // 3)) for x in seqs]
        al = _aligner(params)
        try:
            al.set_option("p16", 1)
            s16 = al.align(seqs, structs, pairs, want_trace=False)
            assert al.engine.stats()["kernel_kind"] == 5
            al.set_option("p16", 0)
            s32 = al.align(seqs, structs, pairs, want_trace=False)
            assert al.engine.stats()["kernel_kind"] in (1, 2, 10)
        finally:
            _unselect(al)
        assert (s16 == s32).all()
        for q, (ia, ib) in enumerate(pairs):
            assert int(s16[q]) == oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes")["score"], q


def test_16bit_pair_mode_refused_when_range_does_not_fit():
    from bialign_b200 import _capi
    from bialign_b200 import workloads
    from bialign_b200.batch import BatchAligner

    res, cls, off, pa, pb = workloads.protein_pairs(4, seed=3)  # length 200-500 with BLOSUM62 x100: needs > 16 bits
    al = BatchAligner(max_shift=2, **workloads.PROTEIN_PARAMS)
    try:
        al.set_option("p16", 1)
        with pytest.raises(_capi.BialignError) as ei:
            al.align_encoded(res, cls, off, pa, pb, want_trace=False)
        assert ei.value.code == _capi.BA_ERR_SCORE_RANGE
    finally:
        _unselect(al)
    al.align_encoded(res, cls, off, pa, pb, want_trace=False)
    assert al.engine.stats()["kernel_kind"] in (1, 3)  # auto: falls back to a 32-bit kernel


def test_wide_band_and_large_alphabet_fall_back_to_the_general_kernel():
    """max_shift 5 (beyond the systolic instantiations) and a 256-symbol similarity table (raw bytes, too large
    for the shared-memory copy of the fast kernel) run on the general level kernel -- still on the GPU, still exact."""
    from bialign_b200 import _capi
    from bialign_b200.batch import trace_hex

    rng = np.random.default_rng(55)
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=5)
    seqs, structs, pairs = _random_protein_batch(rng, 4, 3, 24)
    al = _aligner(params)
    scores, cols, offsets, complete = al.align(seqs, structs, pairs, want_trace=True)
    assert al.engine.stats()["kernel_kind"] == 0
    for q, (ia, ib) in enumerate(pairs):
        r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes")
        assert int(scores[q]) == r["score"] and trace_hex(cols, offsets, q) == r["trace"], q
    # raw-byte residues with the oracle's own 256 x 256 table, straight through the C ABI wrapper
    params["max_shift"] = 2
    eng = _capi.get_engine()
    eng.set_scoring(oracle.blosum62_table(), 800, -150, -50, -150, 2)
    res = np.concatenate([oracle.encode(x, y, False)[0] for x, y in zip(seqs, structs)])
    cls = np.concatenate([oracle.encode(x, y, False)[1] for x, y in zip(seqs, structs)])
    off = np.concatenate([[0], np.cumsum([len(x) for x in seqs])]).astype(np.int64)
    pa = np.array([p[0] for p in pairs], dtype=np.int32)
    pb = np.array([p[1] for p in pairs], dtype=np.int32)
    sc = eng.align_batch(res, cls, off, pa, pb, want_trace=True)
    assert eng.stats()["kernel_kind"] == 0
    cols, offsets, complete = eng.fetch_traces()
    for q, (ia, ib) in enumerate(pairs):
        r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes")
        assert int(sc[q]) == r["score"] and trace_hex(cols, offsets, q) == r["trace"], q


@pytest.mark.parametrize("variant", [{"shift_cost": 0}, {"structure_weight": 0, "gap_cost": 0}, {"shift_cost": 0, "gap_cost": 0, "gap_opening_cost": -1}])
def test_tie_storms_at_multipass_sizes(variant):
    """Tie-breaking exactness where it is hardest: parameter sets that make most cases tie, on pairs long enough for
    a dozen passes through the boundary streams; scores, traces and every reachable code vs the oracle."""
    rng = np.random.default_rng(31337)
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=2)
    params.update(variant)
    seqs, structs, pairs = _random_protein_batch(rng, 3, 150, 260)
    # make the second molecule of each pair a noisy copy of the first: long runs of equally good alternatives
    for q in range(3):
        a = seqs[2 * q]
        b = list(a[: len(seqs[2 * q + 1])].ljust(len(seqs[2 * q + 1]), "A"))
        for p in range(0, len(b), 9):
            b[p] = "ARNDCQEGHILKMFPSTWYV"[rng.integers(0, 20)]
        seqs[2 * q + 1] = "".join(b)
        structs[2 * q + 1] = (structs[2 * q][3:] + "HHH")[: len(b)].ljust(len(b), "C")
    al = _aligner(params)
    assert _check_batch(al, seqs, structs, pairs, params, table_pairs=3) in (3, 4)  # three long-ish pairs: long-pair mode
    al.set_option("long", 0)
    try:
        assert _check_batch(al, seqs, structs, pairs, params, table_pairs=3) in (1, 2)  # CTA-per-pair mode
    finally:
        _unselect(al)


@pytest.mark.parametrize("kernel", [0, 1, 2, 3])
@pytest.mark.parametrize("s", [0, 1, 2, 3, 4])
def test_nonaffine_model_vs_oracle(s, kernel):
    """gap_opening_cost == 0 (the CLI default): level kernel (0), both non-affine flavours of the systolic kernel (1, 2) and
    the dedicated row-per-lane kernel (3, max_shift <= 3) against the oracle's literal restatement of pyx:443-471 /
    513-531 -- scores and first-case-wins traces, ragged multi-block batch."""
    from bialign_b200.batch import trace_hex

    rng = np.random.default_rng(2300 + s)
    for var in ({}, {"shift_cost": 0}, {"gap_cost": 0, "structure_weight": 0}):
        params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=0, gap_cost=-200,
                      shift_cost=-250, max_shift=s)
        params.update(var)
        seqs, structs, pairs = _random_protein_batch(rng, 8, 1, 110)
        al = _aligner(params)
        _select(al, min(kernel, 1) if kernel == 3 else kernel)
        al.set_option("na_kernel", 1 if kernel == 3 and s <= 3 else 0)
        if kernel == 3:
            al.set_option("warps_per_cta", int(rng.choice([1, 2, 3])))  # 32, 64, 96 rows per block: one to four blocks per pair
        try:
            scores, cols, offsets, complete = al.align(seqs, structs, pairs, want_trace=True)
            kind = al.engine.stats()["kernel_kind"]
            assert kind == (0 if kernel == 0 else (8 if s <= 3 else 6) if kernel == 3 else 5 + kernel)
            for q, (ia, ib) in enumerate(pairs):
                r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params)
                assert int(scores[q]) == r["score"], (q, var)
                assert trace_hex(cols, offsets, q) == r["trace"], (q, var)
            assert (al.align(seqs, structs, pairs, want_trace=False) == scores).all()
        finally:
            _unselect(al)


def test_many_waves_with_a_tiny_code_arena():
    """A traceback-code arena far smaller than the batch forces many waves (fill + traceback per wave); results must
    not depend on the wave structure."""
    from bialign_b200.batch import trace_hex

    rng = np.random.default_rng(4242)
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=2)
    seqs, structs, pairs = _random_protein_batch(rng, 40, 40, 120)
    al = _aligner(params)
    ref_scores, ref_cols, ref_off, _ = al.align(seqs, structs, pairs, want_trace=True)
    assert al.engine.stats()["waves"] == 1
    try:
        for kernel in (1, 0):
            al.set_option("kernel", kernel)
            al.set_option("code_arena_bytes", 3 << 20)  # ~ 2-3 pairs per wave
            scores, cols, offsets, complete = al.align(seqs, structs, pairs, want_trace=True)
            assert al.engine.stats()["waves"] >= 10
            assert (scores == ref_scores).all() and complete.all()
            assert all(trace_hex(cols, offsets, q) == trace_hex(ref_cols, ref_off, q) for q in range(len(pairs)))
    finally:
        al.set_option("code_arena_bytes", 0)
        _unselect(al)
    for q in (0, 7, 39):
        ia, ib = pairs[q]
        r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes")
        assert int(ref_scores[q]) == r["score"] and trace_hex(ref_cols, ref_off, q) == r["trace"]


def test_fuzz_slice_vs_literal_oracle():
    """100 rounds of scripts/fuzz_parity.py (random parameters incl. zero / positive costs, RNA and protein, max_shift 0..4,
    every kernel / flavour / CTA width / long-pair mode) against the oracle's literal int64 restatement."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(root, "scripts", "fuzz_parity.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    bad, checked = fz.fuzz(seed=20261018, rounds=100, verbose=False)
    assert checked > 200 and bad == 0


def test_very_long_molecule_b_falls_back_instead_of_narrowing_the_cta():
    """Molecule B is staged in shared memory by the systolic kernel; when it no longer fits next to the rings at the minimum
    CTA width (one boundary-record element per thread) the engine must take the general level kernel -- never a narrower
    CTA, which would drop boundary elements between the row blocks of a multi-pass pair."""
    rng = np.random.default_rng(5)
    aa = "ARNDCQEGHILKMFPSTWYV"
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                  shift_cost=-150, max_shift=1)
    seqs, structs = [], []
    for L in (150, 78000):  # 150 rows = several row blocks at any CTA width
        seqs.append("".join(aa[i] for i in rng.integers(0, 20, L)))
        structs.append("".join("HEC"[i] for i in rng.integers(0, 3, L)))
    from bialign_b200.batch import BatchAligner

    al = BatchAligner(**params)
    s_auto, cols, offsets, complete = al.align(seqs, structs, [(0, 1)], want_trace=True)
    kind_auto = al.engine.stats()["kernel_kind"]
    t_auto = trace_hex_(cols, offsets, 0)
    al.set_option("kernel", 0)
    s_gen, cols, offsets, _ = al.align(seqs, structs, [(0, 1)], want_trace=True)
    assert al.engine.stats()["kernel_kind"] == 0
    assert int(s_auto[0]) == int(s_gen[0]) and t_auto == trace_hex_(cols, offsets, 0) and bool(complete[0])
    v, end = oracle.eval_trace(seqs[0], seqs[1], structs[0], structs[1], params, t_auto)
    assert v == int(s_auto[0]) and end == [150, 78000, 150, 78000]
    assert kind_auto in (0, 1, 2, 3, 4, 11, 12)


def trace_hex_(cols, offsets, p):
    from bialign_b200.batch import trace_hex

    return trace_hex(cols, offsets, p)


def test_probabilistic_structure_similarity_matches_reference(monkeypatch):
    """No structure supplied for an RNA: base-pair probabilities come from the `RNA` module (here tests/fake_rna, the same
    stand-in the goldens were generated with from the unmodified reference), mu2 = int(w * (sqrt(upA upB) + ...)) is
    evaluated on the host and uploaded per pair (ba_set_pair_mu2), the level kernel fills.  Scores, traces, the 14 decoded
    rows (MEA consensus structures of the probability matrices) and eval_trace must equal the reference's."""
    import json
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(root, "tests", "fake_rna"))
    monkeypatch.delitem(sys.modules, "RNA", raising=False)
    from bialign_b200 import bialignment as ba

    cases = json.load(open(os.path.join(root, "tests", "golden", "rna_prob_cases.json")))
    assert len(cases) >= 10
    for c in cases:
        b = ba.BiAligner(c["seqA"], c["seqB"], c["strA"], c["strB"], nameA="A", nameB="B", **c["params"])
        assert int(b.optimize()) == c["score"], (c["seqA"], c["params"])
        tr = b.traceback()
        assert "".join("%x" % (8 * x[0] + 4 * x[1] + 2 * x[2] + x[3]) for x in tr) == c["trace"]
        assert [[n, r] for n, r in b.decode_trace_full(tr)] == c["full"]
        assert list(b.eval_trace(tr))[-2:] == c["eval_tail"]


@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_chained_short_pairs_vs_oracle(s):
    """Batches of pairs that each fit one row block run as chains through the systolic array (kernel_kind 10): every lane
    moves from pair to pair on its own.  Ragged lengths incl. empty molecules, more pairs than one chain holds, tie storms;
    scores, traces, end values and the code table of every reachable cell-state against the oracle, and against chain = 0."""
    rng = np.random.default_rng(4100 + s)
    hi = {0: 30, 1: 60, 2: 40, 3: 28}[s]
    for var in ({}, {"shift_cost": 0, "gap_cost": 0}, {"structure_weight": 0}):
        params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                      shift_cost=-150, max_shift=s)
        params.update(var)
        seqs, structs, pairs = _random_protein_batch(rng, 70, 1, hi)
        for q in (3, 11):  # empty molecules inside a chain
            seqs[2 * q] = ""
            structs[2 * q] = ""
        seqs[2 * 20 + 1] = ""
        structs[2 * 20 + 1] = ""
        pairs = [p for p in pairs if len(seqs[p[0]]) or len(seqs[p[1]])]  # (both empty: the reference raises, pyx:407)
        al = _aligner(params)
        al.set_option("chain", 1)
        kind = _check_batch(al, seqs, structs, pairs, params, table_pairs=len(pairs))
        assert kind == 10
        chained = al.align(seqs, structs, pairs, want_trace=True)
        al.set_option("chain", 0)
        plain = al.align(seqs, structs, pairs, want_trace=True)
        assert al.engine.stats()["kernel_kind"] in (1, 2)
        for x, y in zip(chained, plain):
            assert (x == y).all()


def _same_results(x, y):
    return all((a == b).all() for a, b in zip(x, y))


@pytest.mark.parametrize("s", [0, 1, 2, 3, 4])
def test_rebased_trace_run_forced_vs_oracle(s):
    """The rebased trace run (values relative to the row maxima of a score-only launch, kernel_kind 11) forced on ordinary
    parameter sets: ragged batches incl. multi-pass pairs and tie storms against the oracle; then with a window so small
    that the walks hit the floor: those pairs are recomputed by the level kernel and every answer is still the oracle's."""
    rng = np.random.default_rng(5200 + s)
    for var in ({}, {"shift_cost": 0, "gap_cost": 0}, {"structure_weight": 0, "gap_opening_cost": -1}):
        params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150, gap_cost=-50,
                      shift_cost=-150, max_shift=s)
        params.update(var)
        seqs, structs, pairs = _random_protein_batch(rng, 10, 1, 110)
        seqs[4], structs[4] = "", ""  # an empty molecule A
        al = _aligner(params)
        al.set_option("rebase", 1)
        try:
            assert _check_batch(al, seqs, structs, pairs, params) == 11
            assert al.trace_run_stats["fallback_pairs"] == 0 or var
            al.set_option("rebase_window", 40)
            assert _check_batch(al, seqs, structs, pairs, params) == 11
            assert al.trace_run_stats["fallback_pairs"] > 0
        finally:
            _unselect(al)


def test_rebased_trace_run_beyond_packed_range():
    """Scores without a common divisor (structure_weight 333): value << tie bits leaves 32 bits beyond ~800 residues (even for
    the padded flavour).  The engine then picks the rebased run on its own (kernel_kind 11, or 12 when the pairs run as long-pair gangs) instead of the level kernel.
    Oracle on a subsample, the level kernel (kernel = 0) on everything: scores, traces, completeness."""
    from bialign_b200.batch import trace_hex

    rng = np.random.default_rng(77001)
    params = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=333, gap_opening_cost=-157, gap_cost=-49,
                  shift_cost=-151, max_shift=2)
    seqs, structs, pairs = _random_protein_batch(rng, 22, 850, 1000)
    sq2, st2, _ = _random_protein_batch(rng, 2, 250, 350)  # two shorter pairs in the same batch: the oracle's share
    pairs += [(len(seqs), len(seqs) + 1), (len(seqs) + 2, len(seqs) + 3)]
    seqs, structs = seqs + sq2, structs + st2
    al = _aligner(params)
    fast = al.align(seqs, structs, pairs, want_trace=True)
    st = al.engine.stats()
    assert st["kernel_kind"] in (11, 12) and st["fallback_pairs"] <= 2, st  # 12: few enough pairs for long-pair gangs
    al.set_option("kernel", 0)
    try:
        slow = al.align(seqs, structs, pairs, want_trace=True)
        assert al.engine.stats()["kernel_kind"] == 0
    finally:
        _unselect(al)
    assert _same_results(fast, slow)
    scores, cols, offsets, complete = fast
    for q in (22, 23):
        ia, ib = pairs[q]
        r = oracle.run(seqs[ia], seqs[ib], structs[ia], structs[ib], params, mode="codes")
        assert int(scores[q]) == r["score"] and trace_hex(cols, offsets, q) == r["trace"] and bool(complete[q]) == r["complete"]
    # two long pairs: long-pair mode, several row blocks per CTA
    aa = "ARNDCQEGHILKMFPSTWYV"
    seqs, structs = [], []
    for L in (1900, 2000, 2100, 1800):
        seqs.append("".join(aa[i] for i in rng.integers(0, 20, L)))
        structs.append("".join("HEC"[i] * 7 for i in rng.integers(0, 3, L // 7 + 1))[:L])
    params["max_shift"] = 1
    al = _aligner(params)
    fast = al.align(seqs, structs, [(0, 1), (2, 3)], want_trace=True)
    st = al.engine.stats()
    assert st["kernel_kind"] == 12 and st["fallback_pairs"] == 0, st
    al.set_option("kernel", 0)
    try:
        slow = al.align(seqs, structs, [(0, 1), (2, 3)], want_trace=True)
    finally:
        _unselect(al)
    assert _same_results(fast, slow)
