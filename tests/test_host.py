"""CPU tests of the host side: C-ABI export list, encoders, scheduler arithmetic, presentation layer
(against rows recorded from the unmodified reference).  No GPU, no compute calls into the library."""
import io
import contextlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from bialign_b200 import _capi

    header = open(os.path.join(ROOT, "include", "bialign_b200.h")).read()
    declared = set(re.findall(r"BA_API[^;(]*?\b(ba_[a-z_0-9]+)\s*\(", header))
    assert declared and declared == set(_capi.SYMBOLS)
    lib = _capi.load_library()
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.ba_version()


def test_no_cpu_path_without_device():
    """The product must fail loudly without a GPU (this container has none; on the GPU box this is skipped)."""
    from bialign_b200 import _capi

    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(_capi.BialignError) as ei:
        _capi.Engine(0)
    assert ei.value.code == _capi.BA_ERR_NO_DEVICE
    from bialign_b200 import bialignment as ba

    b = ba.BiAligner("AC", "AC", "HH", "HH", type="Protein", gap_cost=-50, gap_opening_cost=-150, max_shift=1,
                     simmatrix="BLOSUM62", structure_weight=800, shift_cost=-150)
    with pytest.raises(_capi.BialignError):
        b.optimize()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bialign_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "bialign_oracle" not in txt and "oracle/" not in txt, f


def test_band_cells_formula_against_enumeration():
    from bialign_b200.batch import band_cells

    for n, m, s in [(0, 0, 1), (1, 3, 2), (5, 5, 0), (7, 4, 3), (3, 9, 4), (42, 42, 1)]:
        cnt = 0
        for i in range(n + 1):
            for j in range(m + 1):
                cnt += (min(n, i + s) - max(0, i - s) + 1) * (min(m, j + s) - max(0, j - s) + 1)
        assert int(band_cells(n, m, s)) == cnt
    assert int(band_cells(42, 42, 1)) * 9 == 145161  # config 1


def test_lpt_shards_partition_and_balance():
    from bialign_b200.batch import lpt_shards

    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 100, 5000):
        costs = rng.integers(1, 1000, n)
        for w in (1, 2, 8):
            sh = lpt_shards(costs, w)
            allidx = np.sort(np.concatenate(sh)) if n else np.zeros(0)
            assert allidx.size == n and (allidx == np.arange(n)).all()
            if n >= 100:
                loads = np.array([costs[x].sum() for x in sh])
                assert loads.max() <= loads.mean() * 1.05 + costs.max()


def test_encoders_match_oracle_side():
    import oracle
    from bialign_b200 import encoding

    t_o = oracle.blosum62_table()
    m = encoding.read_simmatrix("BLOSUM62")
    for a in encoding.ALPHABET:
        for b in encoding.ALPHABET:
            assert m[a][b] == t_o[ord(a), ord(b)]
    for st in ["...(((.....))).....", "()", "(())..((", ".(.).", "((.))()", "", "(((", "a(b)c"]:
        assert (encoding.rna_structure_classes(st) == oracle.rna_classes(st)).all()
    with pytest.raises(IndexError):
        encoding.rna_structure_classes("())")  # unbalanced ')' like pyx:387


def test_constructor_error_conventions():
    from bialign_b200 import bialignment as ba

    base = dict(type="Protein", gap_cost=-50, gap_opening_cost=-150, max_shift=1, simmatrix=None)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), pytest.raises(SystemExit) as ei:
        ba.BiAligner("ACD", "ACD", None, "HHH", **base)
    assert ei.value.code == -1 and buf.getvalue().startswith("ERROR: Structures have to be provided")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), pytest.raises(SystemExit):
        ba.BiAligner("ACD", "ACD", "HH", "HHH", **base)
    assert "must have the same length" in buf.getvalue()
    with pytest.raises(KeyError):
        ba.BiAligner("ACD", "ACD", "HHH", "HHH", type="Protein", gap_cost=-50, max_shift=1, simmatrix=None)


def _aligner_for(case):
    from bialign_b200 import bialignment as ba

    return ba.BiAligner(case["seqA"], case["seqB"], case["strA"], case["strB"], nameA="A", nameB="B", **case["params"])


def test_decode_trace_full_matches_reference_rows(golden_cases):
    """The 14 named rows (incl. the RNA consensus structure via the MEA fold) for every golden trace."""
    checked = 0
    for c in golden_cases:
        b = _aligner_for(c)
        trace = [[int(ch, 16) >> 3 & 1, int(ch, 16) >> 2 & 1, int(ch, 16) >> 1 & 1, int(ch, 16) & 1] for ch in c["trace"]]
        if c["params"]["gap_opening_cost"] == 0:
            trace = [tuple(x) for x in trace]
        got = [[n, r] for n, r in b.decode_trace_full(trace)]
        assert got == c["full"], (c["seqA"], c["seqB"], c["params"])
        ev = list(b.eval_trace(trace))
        assert ev[-2:] == c["eval_tail"]
        checked += 1
    assert checked == len(golden_cases)


def test_outmodes_and_readme_rna_default_output(golden_cases):
    from bialign_b200 import bialignment as ba

    c = golden_cases[0]  # README RNA toy (README.md:90-104)
    b = _aligner_for(c)
    trace = [[int(ch, 16) >> 3 & 1, int(ch, 16) >> 2 & 1, int(ch, 16) >> 1 & 1, int(ch, 16) & 1] for ch in c["trace"]]
    lines = b.decode_trace(trace)
    assert lines == ["A               GCGGGGGAUAUCCCC-AUCG", "B               G---GGGAUAUCCCC-AUCG",
                     "A ss            ...-(((.....))).....", "B ss            .---(((.....)))-....",
                     "A shifts        ...<...........>....", "B shifts        ...................."]
    assert ba.BiAligner.auto_complete("sorted_t", ba.BiAligner.outmodes.keys()) == "sorted_terse"
    assert ba.BiAligner.auto_complete("s", ba.BiAligner.outmodes.keys()) == "sorted"
    p = golden_cases[3]  # README protein toy, max_shift 1, sorted mode (README.md:136-152)
    b = _aligner_for(p)
    b._params["outmode"] = "sorted"
    trace = [[int(ch, 16) >> 3 & 1, int(ch, 16) >> 2 & 1, int(ch, 16) >> 1 & 1, int(ch, 16) & 1] for ch in p["trace"]]
    out = b.decode_trace(trace)
    assert out[0] == "A ss            -CHHHHHHHHHHHHHCCCCTCEEEEEEECCTCEEEEEEEEC-CC"
    assert out[2] == "consensus       -.AKLPLKEKKLT.TANYHPGIRYIMTGYSAK.IYSSTYA.-FR"
    assert out[6] == "" and out[-2] == "A shifts        >............<..................<........>.."
    b._params["outmode"] = "nonsense"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out2 = b.decode_trace(trace)
    assert buf.getvalue().startswith("WARNING: unknown output mode") and out2 == out


def test_cfssp_reader_on_synthetic_report(tmp_path):
    from bialign_b200 import bialignment as ba

    txt = "Secondary Structure:\n\nQuery 1   MVQIP 5 \nHelix 1   HHH   5 \nStruc 1   EEHHC 5 \n\nQuery 6   AK 7 \nStruc 6   CC 7 \n"
    assert ba.read_molecule(txt, "Protein") == ["MVQIPAK", "EEHHCCC"]
    with pytest.raises(IOError):
        ba.read_molecule(txt, "RNA")
    f = tmp_path / "x.cfssp"
    f.write_text(txt)
    assert ba.read_molecule_from_file(str(f), "Protein") == ["MVQIPAK", "EEHHCCC"]


def test_read_molecule_from_file_error_paths(tmp_path, capsys):
    """Messages of nonpyx:80-93 (compared with the compiled reference when this was written: same text on stdout; the
    reference then dies with a NameError because its module never imports sys -- here the intended exit status -1)."""
    from bialign_b200 import bialignment as ba

    f = tmp_path / "x.cfssp"
    f.write_text("Query 1   MV 2 \nStruc 1   EE 2 \n")
    cases = [(str(tmp_path / "missing"), "Protein", "Input file not found.\n[Errno 2] No such file or directory: '%s'\n" % (tmp_path / "missing")),
             (str(tmp_path), "Protein", "Cannot read input file %s.\n[Errno 21] Is a directory: '%s'\n" % (tmp_path, tmp_path)),
             (str(f), "RNA", "Cannot read input file %s.\nCannot read files of type RNA\n" % f)]
    for name, kind, text in cases:
        with pytest.raises(SystemExit) as ex:
            ba.read_molecule_from_file(name, kind)
        assert ex.value.code == -1
        assert capsys.readouterr().out == text


def test_compact_shard_keeps_sequences():
    from bialign_b200 import workloads
    from bialign_b200.batch import compact_shard

    res, cls, off, pa, pb = workloads.protein_pairs(50, lo=5, hi=20, seed=1)
    mine = np.array([3, 7, 8, 20, 49])
    r2, c2, o2, a2, b2 = compact_shard(res, cls, off, pa[mine], pb[mine])
    assert len(r2) == len(c2) == o2[-1] < len(res)
    for q, p in enumerate(mine):
        assert (r2[o2[a2[q]]:o2[a2[q] + 1]] == res[off[pa[p]]:off[pa[p] + 1]]).all()
        assert (c2[o2[b2[q]]:o2[b2[q] + 1]] == cls[off[pb[p]]:off[pb[p] + 1]]).all()


def test_workload_generators_are_deterministic_and_well_formed():
    from bialign_b200 import encoding, workloads

    r1 = workloads.protein_pairs(20, seed=3)
    r2 = workloads.protein_pairs(20, seed=3)
    assert all((x == y).all() for x, y in zip(r1, r2))
    lens = np.diff(r1[2])
    assert lens.min() >= 200 and lens.max() <= 500 and set(bytes(r1[1]).decode()) <= set("HEC")
    res, cls, off, pa, pb = workloads.rna_pairs(8, seed=4)
    for q in range(16):
        seq, st = workloads.decode_rna(res, cls, off, q)
        assert len(seq) == 120 and st.count("(") == st.count(")")
        assert (encoding.rna_structure_classes(st) == cls[off[q]:off[q + 1]]).all()


def test_header_is_valid_c_and_links(tmp_path):
    """include/bialign_b200.h is plain C (no C++/CUDA types); a C client compiles and links against the library."""
    import subprocess

    from bialign_b200 import _capi

    src = tmp_path / "client.c"
    src.write_text('#include "bialign_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { ba_engine* e = 0; int rc = ba_engine_create(0, &e);\n'
                   '  printf("%s rc=%d %s\\n", ba_version(), rc, rc ? ba_last_error(0) : "ok");\n'
                   '  if (!rc) { ba_engine_destroy(e); }\n  return 0; }\n')
    exe = tmp_path / "client"
    libdir = os.path.dirname(_capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libbialign_b200.so", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "bialign_b200" in out.stdout


def test_missing_library_fails_loudly(monkeypatch):
    """No fallback: without the built CUDA library the loader raises (and says how to build it)."""
    from bialign_b200 import _capi

    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libbialign_b200.so")
    with pytest.raises(ImportError) as ei:
        _capi.load_library()
    assert "no CPU implementation" in str(ei.value)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's own CPU path from oracle/_ref, or the C port): one JSON line with the
    contract keys on rank 0, nothing (exit 0) on the other ranks of a torchrun launch.  No GPU involved."""
    import json
    import subprocess
    import sys

    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--pairs-per-gpu", "64"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    other = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, env=env, timeout=300)
    assert other.returncode == 0 and other.stdout.strip() == ""
    # the GPU count on the command line must agree with the launcher's world size
    bad = subprocess.run(cmd + ["--gpus", "4"], capture_output=True, text=True, env=env, timeout=300)
    assert bad.returncode != 0 and "WORLD_SIZE" in bad.stderr
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-400:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GCUPS" and line["higher_is_better"] is True
    assert line["metric"] == "batched bialign GCUPS (cell-states/s)" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] and "sample" in line["cpu_baseline"]
    assert line["e2e"] == {"value": line["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("cfg3") and line["vs_baseline"] is None
    assert line["config"]["reference_sample"]["pairs"] >= 1  # the bounded CPU sample is declared in the arm's own config


def _write_matrix(path, symbols, rows, trailer=()):
    with open(path, "w") as fh:
        fh.write("-  " + "  ".join(symbols) + "\n")
        for s, r in zip(symbols, rows):
            fh.write(s + " " + " ".join("%2d" % v for v in r) + " \n")
        for t in trailer:
            fh.write(t + "\n")


def test_read_simmatrix_file_format(tmp_path, capsys):
    """File branch of read_simmatrix (nonpyx:33-58): header row starting with '-', one row per symbol, entries x 100,
    anything after the last row ignored; a row/column order mismatch prints the reference's (literal) error line."""
    from bialign_b200 import encoding

    builtin = encoding.read_simmatrix("BLOSUM62")
    syms = list(encoding.ALPHABET)
    f = tmp_path / "b62.txt"
    _write_matrix(f, syms, encoding.blosum62_rows(), trailer=["this line and the next are never parsed", "X Y Z"])
    assert encoding.read_simmatrix(str(f)) == builtin
    assert encoding.read_simmatrix(str(f), scale=1)["W"]["W"] == 11
    assert capsys.readouterr().out == ""
    ref_file = "/root/reference/Data/BLOSUM62.txt"
    if os.path.exists(ref_file):  # build container only: the reference's own data file holds the same numbers
        assert encoding.read_simmatrix(ref_file) == builtin
    g = tmp_path / "swapped.txt"
    _write_matrix(g, ["A", "C"], [[1, 2], [3, 4]])
    text = open(g).read().replace("\nA ", "\nQ ", 1)
    open(g, "w").write(text)
    m = encoding.read_simmatrix(str(g))
    assert m == {"Q": {"A": 100, "C": 200}, "C": {"A": 300, "C": 400}}
    assert capsys.readouterr().out == "ERROR while reading simmatrix {filename}.\n"


def test_batch_aligner_rejects_unknown_arguments_and_undefined_matrix_entries(tmp_path):
    """A misspelled scoring argument must not silently run with defaults, and a residue pair the matrix does not define
    raises KeyError like the reference's dict lookup (pyx:407) instead of scoring 0."""
    from bialign_b200.batch import BatchAligner

    with pytest.raises(TypeError):
        BatchAligner(type="Protein", simmatrix="BLOSUM62", gap_costs=-50)
    f = tmp_path / "holes.txt"
    with open(f, "w") as fh:  # three columns, and the row of C is short: (C, D) is undefined
        fh.write("- A C D\nA 4 0 1\nC 0 9\nD 1 2 6\n")
    al = BatchAligner(type="Protein", simmatrix=str(f), gap_opening_cost=-10)
    assert al.known is not None and not al.known.all()
    res, cls, off = al.encode(["ADA", "AAD", "CC"], ["HHH", "HHH", "HH"])
    al.check_known(res, off, [0], [1])  # A/D rows against A/D columns: all defined
    with pytest.raises(KeyError):
        al.check_known(res, off, [2], [1])  # C against D


def test_probabilistic_profile_host_arithmetic(monkeypatch):
    """Host side of the probabilistic RNA path (no GPU): the profile of a predicted structure (tests/fake_rna stands in for
    ViennaRNA), the uploaded integer matrix and the scalar mu2() agree, and a supplied structure gives the one-hot profile."""
    import sys

    monkeypatch.syspath_prepend(os.path.join(ROOT, "tests", "fake_rna"))
    monkeypatch.delitem(sys.modules, "RNA", raising=False)
    from bialign_b200 import bialignment as ba

    params = dict(type="RNA", simmatrix=None, structure_weight=333, gap_opening_cost=-200, gap_cost=-50, shift_cost=-150,
                  max_shift=1, sequence_match_similarity=100, sequence_mismatch_similarity=0)
    b = ba.BiAligner("GGGAAAUCCCGAUUAGCUAGC", "GGCAUAUGCCGAUCG", None, "((..)).........", **params)
    assert b.molA.get("predicted") and not b.molB.get("predicted")
    n, m = b.molA["len"], b.molB["len"]
    for key in ("up", "down", "unp"):
        assert len(b.molA[key]) == n + 1 and len(b.molB[key]) == m + 1
    assert all(abs(u + d + p - 1.0) < 1e-12 for u, d, p in zip(b.molA["up"], b.molA["down"], b.molA["unp"]))
    assert b.molB["down"][1:3] == [1.0, 1.0] and b.molB["up"][5:7] == [1.0, 1.0] and b.molB["unp"][0] == 1.0
    mat = b._mu2_matrix()
    assert mat.shape == (n, m) and mat.dtype == np.int32
    assert all(mat[k - 1, l - 1] == b.mu2(k, l) for k in range(1, n + 1) for l in range(1, m + 1))
    assert 0 < mat.max() <= 333 and (mat % 333 != 0).any()  # genuinely fractional similarities, truncated
    sb = b.molA["sbpp"]
    assert np.allclose(sb, sb.T) and np.allclose(sb[1:, 1:].sum(axis=1), 1.0)
    # supplied structures only: no matrix, the class path is used
    assert ba.BiAligner("GGGAAACCC", "GGAAACC", "(((...)))", "((...))", **params)._mu2_matrix() is None


def test_rna_without_structure_needs_the_rna_module(monkeypatch):
    """Like the reference (pyx:347), an RNA without a supplied structure imports ViennaRNA's `RNA` module and fails with
    ImportError when it is not installed; proteins without structures print the reference's error and exit."""
    import sys

    monkeypatch.setitem(sys.modules, "RNA", None)
    from bialign_b200 import bialignment as ba

    params = dict(type="RNA", simmatrix=None, structure_weight=400, gap_opening_cost=-200, gap_cost=-50, shift_cost=-150,
                  max_shift=1, sequence_match_similarity=100, sequence_mismatch_similarity=0)
    with pytest.raises(ImportError):
        ba.BiAligner("GGGAAACCC", "GGAAACC", None, None, **params)


def test_every_engine_option_is_documented_and_known_to_the_library():
    """The Python layer's option table, the header's ba_set_option comment and the library's own key list agree
    (no compute call: the keys are looked up in the strings of the built library)."""
    from bialign_b200 import _capi

    header = open(os.path.join(ROOT, "include", "bialign_b200.h")).read()
    lib = open(os.path.join(ROOT, "bialign_b200", "libbialign_b200.so"), "rb").read()
    for key in _capi.ENGINE_OPTIONS:
        assert '"%s"' % key in header, key
        assert key.encode() + b"\0" in lib, key
