"""Synthetic workloads of BASELINE.json / SURVEY 8d (numpy default_rng, deterministic).

cfg3: protein pairs, lengths U{200..500}, residues uniform over the 20 amino acids, structure =
      concatenated runs (symbol uniform in H,E,C; run length U{3..12}), README protein scoring, max_shift 2.
cfg4: RNA pairs of length 120, residues uniform ACGU, balanced dot-bracket from a random stack
      process, README RNA scoring, max_shift 2, score only.
cfg5: one protein pair 8192 x 8192, max_shift 3.
All generators return encoded arrays for the C ABI: (residues, classes, offsets, pair_a, pair_b).
"""
import numpy as np

from . import encoding

AA20 = "ARNDCQEGHILKMFPSTWYV"
PROTEIN_PARAMS = dict(type="Protein", simmatrix="BLOSUM62", structure_weight=800, gap_opening_cost=-150,
                      gap_cost=-50, shift_cost=-150)
RNA_PARAMS = dict(type="RNA", simmatrix=None, structure_weight=400, gap_opening_cost=-200, gap_cost=-50,
                  shift_cost=-150, sequence_match_similarity=100, sequence_mismatch_similarity=0)


def _run_structures(rng, lens):
    """One H/E/C run structure per sequence, vectorised: draw enough runs for the longest sequence,
    expand, keep the first len[q] symbols of sequence q."""
    lens = np.asarray(lens, dtype=np.int64)
    nseq = lens.size
    if nseq == 0:
        return np.zeros(0, dtype=np.uint8)
    K = int(lens.max()) // 3 + 2
    rl = rng.integers(3, 13, size=(nseq, K))
    sym = rng.integers(0, 3, size=(nseq, K))
    full = np.repeat(np.frombuffer(b"HEC", dtype=np.uint8)[sym.ravel()], rl.ravel())
    row_start = np.concatenate([[0], np.cumsum(rl.sum(axis=1))[:-1]])
    out_start = np.concatenate([[0], np.cumsum(lens)[:-1]])
    idx = np.arange(int(lens.sum())) - np.repeat(out_start, lens) + np.repeat(row_start, lens)
    return full[idx]


def protein_pairs(npairs, lo=200, hi=500, seed=3):
    """cfg3-shaped batch.  Returns (res, cls, off, pair_a, pair_b, symbols): residues are codes over
    the BLOSUM62 alphabet order of encoding.ALPHABET, classes are the raw structure bytes."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, size=(npairs, 2)).reshape(-1)
    total = int(lens.sum())
    aa_codes = np.array([encoding.ALPHABET.index(ch) for ch in AA20], dtype=np.uint8)
    res = aa_codes[rng.integers(0, 20, size=total)]
    cls = _run_structures(rng, lens)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pa = np.arange(0, 2 * npairs, 2, dtype=np.int32)
    pb = pa + 1
    return res, cls, off, pa, pb


def rna_pairs(npairs, length=120, seed=4):
    """cfg4-shaped batch: residues 0..3 = ACGU; classes = up/down/unpaired of a random balanced structure."""
    rng = np.random.default_rng(seed)
    nseq = 2 * npairs
    res = rng.integers(0, 4, size=nseq * length).astype(np.uint8)
    u = rng.random(size=(nseq, length))
    # stack process per sequence (open p=.3; close p=.3 if the innermost open is >= 3 back; else '.'), run for all
    # sequences in lock step: one vectorised update per position
    partner = np.full((nseq, length), -1, dtype=np.int32)
    stack = np.zeros((nseq, length), dtype=np.int32)
    sp = np.zeros(nseq, dtype=np.int32)
    rows = np.arange(nseq)
    for i in range(length):
        ui = u[:, i]
        opening = ui < 0.3
        top = stack[rows, np.maximum(sp - 1, 0)]
        closing = (~opening) & (ui < 0.6) & (sp > 0) & (i - top >= 3)
        ro = rows[opening]
        stack[ro, sp[ro]] = i
        rc = rows[closing]
        jc = top[closing]
        partner[rc, i] = jc
        partner[rc, jc] = i
        sp += opening.astype(np.int32) - closing.astype(np.int32)
    pos = np.arange(length, dtype=np.int32)[None, :]
    paired = partner >= 0
    cls = np.zeros((nseq, length), dtype=np.uint8)
    cls[paired & (partner <= pos - 2)] = encoding.UP
    cls[paired & (partner >= pos + 1)] = encoding.DOWN
    off = (np.arange(nseq + 1) * length).astype(np.int64)
    pa = np.arange(0, nseq, 2, dtype=np.int32)
    return res, cls.reshape(-1), off, pa, pa + 1


def decode_protein(res, cls, off, q):
    """Strings of sequence q (for the CPU baselines / oracle)."""
    a, b = int(off[q]), int(off[q + 1])
    return "".join(encoding.ALPHABET[c] for c in res[a:b]), bytes(cls[a:b]).decode("latin-1")


def decode_rna(res, cls, off, q):
    a, b = int(off[q]), int(off[q + 1])
    seq = "".join("ACGU"[c] for c in res[a:b])
    # any structure string with these classes: rebuild brackets from the class sequence
    st = []
    for c in cls[a:b]:
        st.append("." if c == encoding.UNP else (")" if c == encoding.UP else "("))
    return seq, "".join(st)
