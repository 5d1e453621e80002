"""ctypes front-end of the CPU oracle (oracle/bialign_oracle.c).  TEST INFRASTRUCTURE ONLY.

Encodings here are deliberately independent of bialign_b200's host code so that the product's
encoders are cross-checked too: residues are raw bytes (nsym = 256), the similarity table is a
dense 256x256 int32 matrix built from the reference's own BLOSUM62 text (nonpyx:5-58 semantics:
entries x100) or from match/mismatch (pyx:409-412), structure classes follow pyx:366-392.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

BLOSUM62_ROWS = """-  A  R  N  D  C  Q  E  G  H  I  L  K  M  F  P  S  T  W  Y  V  B  Z  X  *
A  4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4
R -1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4
N -2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4
D -2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4
C  0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4
Q -1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4
E -1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
G  0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4
H -2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4
I -1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4
L -1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4
K -1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4
M -1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4
F -2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4
P -1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4
S  1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4
T  0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4
W -3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4
Y -2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4
V  0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4
B -2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4
Z -1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
X  0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4
* -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_ORACLE_DIR, "libbialign_oracle.so")
        src = os.path.join(_ORACLE_DIR, "bialign_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _ORACLE_DIR, "libbialign_oracle.so"], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(so)
        L.bao_code_words.restype = ctypes.c_size_t
        L.bao_code_index.restype = ctypes.c_size_t
        L.bao_eval_affine_trace.restype = ctypes.c_int64
        _LIB = L
    return _LIB


def blosum62_table():
    t = np.zeros((256, 256), dtype=np.int32)
    lines = BLOSUM62_ROWS.strip("\n").split("\n")
    keys = lines[0].split()[1:]
    for ln in lines[1:]:
        f = ln.split()
        for k, v in zip(keys, f[1:]):
            t[ord(f[0]), ord(k)] = 100 * int(v)
    return t


def match_table(match, mismatch):
    t = np.full((256, 256), mismatch, dtype=np.int32)
    t[np.arange(256), np.arange(256)] = match
    return t


def rna_classes(structure):
    """pyx:366-392 for a supplied dot-bracket string: 0 = unp, 1 = up, 2 = down."""
    n = len(structure)
    partner = [0] * (n + 1)
    stack = []
    for i, c in enumerate(structure):
        if c == "(":
            stack.append(i)
        elif c == ")":
            j = stack.pop()  # IndexError on unbalanced ')', like pyx:387
            partner[i + 1] = j + 1
            partner[j + 1] = i + 1
    cls = np.zeros(n, dtype=np.uint8)
    for i in range(1, n + 1):
        p = partner[i]
        if p == 0:
            continue
        if p <= i - 2:
            cls[i - 1] = 1
        elif p >= i + 1:
            cls[i - 1] = 2
    return cls


def encode(seq, struct, is_rna):
    r = np.frombuffer(seq.encode("latin-1"), dtype=np.uint8).copy()
    if is_rna:
        c = rna_classes(struct)
    else:
        c = np.frombuffer(struct.encode("latin-1"), dtype=np.uint8).copy()
    return r, c


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def sim_for(params):
    if params.get("simmatrix"):
        assert params["simmatrix"] == "BLOSUM62"
        return blosum62_table()
    return match_table(params["sequence_match_similarity"], params["sequence_mismatch_similarity"])


def run(seqA, seqB, strA, strB, params, mode="literal", want_codes=False):
    """Returns dict(score, trace(hex str), complete, codes?)."""
    L = lib()
    is_rna = params["type"] == "RNA"
    ra, ca = encode(seqA, strA, is_rna)
    rb, cb = encode(seqB, strB, is_rna)
    sim = sim_for(params)
    n, m, s = len(ra), len(rb), int(params["max_shift"])
    w, beta, gamma, Delta = (int(params["structure_weight"]), int(params["gap_opening_cost"]),
                             int(params["gap_cost"]), int(params["shift_cost"]))
    score = ctypes.c_int64(0)
    trace = np.zeros(2 * (n + m) + 8, dtype=np.uint8)
    tlen = ctypes.c_int(0)
    comp = ctypes.c_int(1)
    u8, i32 = ctypes.c_uint8, ctypes.c_int32
    out = {}
    if beta == 0:
        rc = L.bao_nonaffine(_p(ra, u8), _p(ca, u8), n, _p(rb, u8), _p(cb, u8), m, _p(sim, i32), 256, w, gamma,
                             Delta, s, ctypes.byref(score), _p(trace, u8), ctypes.byref(tlen))
    elif mode == "literal":
        rc = L.bao_affine_literal(_p(ra, u8), _p(ca, u8), n, _p(rb, u8), _p(cb, u8), m, _p(sim, i32), 256, w, beta,
                                  gamma, Delta, s, ctypes.byref(score), _p(trace, u8), ctypes.byref(tlen),
                                  ctypes.byref(comp))
    else:
        nw = L.bao_code_words(n, m, s)
        codes = np.zeros(nw, dtype=np.uint64)
        endv = np.zeros(9, dtype=np.int32)
        rc = L.bao_affine_codes(_p(ra, u8), _p(ca, u8), n, _p(rb, u8), _p(cb, u8), m, _p(sim, i32), 256, w, beta,
                                gamma, Delta, s, ctypes.byref(score), _p(codes, ctypes.c_uint64), _p(endv, i32),
                                _p(trace, u8), ctypes.byref(tlen), ctypes.byref(comp))
        out["end_values"] = endv
        if want_codes:
            out["codes"] = codes
    assert rc == 0
    out.update(score=int(score.value), trace="".join("%x" % c for c in trace[: tlen.value]), complete=bool(comp.value))
    return out


def eval_trace(seqA, seqB, strA, strB, params, trace_hex):
    L = lib()
    is_rna = params["type"] == "RNA"
    ra, ca = encode(seqA, strA, is_rna)
    rb, cb = encode(seqB, strB, is_rna)
    sim = sim_for(params)
    tr = np.array([int(c, 16) for c in trace_hex], dtype=np.uint8)
    end = np.zeros(4, dtype=np.int32)
    u8, i32 = ctypes.c_uint8, ctypes.c_int32
    v = L.bao_eval_affine_trace(_p(ra, u8), _p(ca, u8), len(ra), _p(rb, u8), _p(cb, u8), len(rb), _p(sim, i32), 256,
                                int(params["structure_weight"]), int(params["gap_opening_cost"]),
                                int(params["gap_cost"]), int(params["shift_cost"]), _p(tr, u8), len(tr), _p(end, i32))
    return int(v), end.tolist()
