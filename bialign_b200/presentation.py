"""Host-side presentation helpers of the drop-in surface (no DP, no GPU): consensus rows, the RNA
consensus structure (maximum-expected-accuracy fold of the consensus pair matrix), CFSSP input
files.  Behavioural spec: src/bialignment.pyx:835-950 and src/bialignment_nonpyx.py:61-112 of the
reference; written from that behaviour, not from its text."""
import sys
from collections import defaultdict

import numpy as np


def consensus_sequence(alistrA, alistrB):
    """Column-wise consensus of two equally long alignment rows: the (upper-cased) symbol where the
    rows agree, '.' elsewhere (pyx:901-908)."""
    return "".join(x if x == y else "." for x, y in zip(alistrA.upper(), alistrB.upper()))


def highlight_sequence_identity(alistrA, alistrB):
    """Lower-case both rows, upper-case the columns where they agree (pyx:890-898)."""
    a, b = [], []
    for x, y in zip(alistrA.lower(), alistrB.lower()):
        if x == y:
            x = y = x.upper()
        a.append(x)
        b.append(y)
    return ["".join(a), "".join(b)]


def parse_dotbracket(dbstr):
    """0-based partner of every position, -1 for unpaired (pyx:911-922)."""
    res = [-1] * len(dbstr)
    stack = []
    for i, sym in enumerate(dbstr):
        if sym == "(":
            stack.append(i)
        elif sym == ")":
            j = stack.pop()
            res[i], res[j] = j, i
    return res


def consensus_sbpp(alistrA, sbppA, alistrB, sbppB):
    """Consensus pair matrix of an alignment: entry (c0, c1) (1-based alignment columns) is
    sqrt(pA * pB) of the residues the two columns hold in A and in B, 0 if either column has a gap
    in that molecule (pyx:926-950).  Vectorised: map columns to residue positions, gather, multiply."""
    L = len(alistrA)
    out = np.zeros((L + 1, len(alistrB) + 1), dtype=float)
    prod = None
    for alistr, sbpp in ((alistrA, sbppA), (alistrB, sbppB)):
        sbpp = np.asarray(sbpp, dtype=float)
        nongap = np.array([c != "-" for c in alistr[:L]], dtype=bool)
        pos = np.cumsum(nongap)  # 1-based residue index held by each non-gap column
        pos_safe = np.where(nongap, pos, 0)
        p = sbpp[np.ix_(pos_safe, pos_safe)] * np.outer(nongap, nongap)
        prod = p if prod is None else prod * p
    out[1:, 1:L + 1] = np.sqrt(prod)
    return out


def mea(sbpp, gamma=3, *, brackets="()"):
    """Maximum-expected-accuracy structure of a symmetric pair matrix with unpaired weights on the
    diagonal; returns (structure string, score).  Same recursion, candidate lists, strict-improvement
    updates and traceback order as pyx:836-886, so the emitted structure is identical -- including
    that reference's quirks (e.g. candidates are only recorded when they strictly improve F)."""
    n = len(sbpp) - 1
    F = np.zeros((n + 2, n + 2), dtype=float)
    T = np.zeros((n + 2, n + 2), dtype=int)
    cands = [[] for _ in range(n + 1)]
    for i in range(n, 0, -1):
        cands[i].append((i, sbpp[i, i]))
        for j in range(i, n + 1):
            best, arg = F[i, j], T[i, j]
            for k, C in cands[j]:
                v = F[i, k - 1] + C
                if best < v:
                    best, arg = v, k
            F[i, j], T[i, j] = best, arg
            if i + 3 >= j:
                continue
            C = F[i + 1, j - 1] + 2 * gamma * sbpp[i, j]
            if C > F[i, j]:
                cands[j].append((i, C))
                F[i, j] = C
                T[i, j] = i
    structure = ["."] * (n + 1)
    stack = [(1, n)]
    while stack:
        i, j = stack.pop()
        if i > n or j < 1:
            continue
        k = T[i, j]
        if i + 3 >= j or k == 0:
            continue
        if k == j:
            stack.append((i, j - 1))
        elif k == i:
            structure[k], structure[j] = brackets[0], brackets[1]
            stack.append((k + 1, j - 1))
        else:
            stack.append((i, k - 1))
            stack.append((k + 1, j - 1))
            structure[k], structure[j] = brackets[0], brackets[1]
    return "".join(structure[1:]), (F[1, n] if n >= 1 else 0.0)


def read_molecule(content, type):
    """Sequence and structure from a CFSSP report: concatenated third fields of the `Query` and
    `Struc` rows (nonpyx:61-82)."""
    if type != "Protein":
        raise IOError(f"Cannot read files of type {type}")
    got = defaultdict(str)
    for line in content.split("\n"):
        f = line.split()
        if f and f[0] in ("Query", "Struc"):
            if len(f) != 4:
                raise IOError("Cannot parse")
            got[f[0]] += f[2]
    if len(got["Query"]) != len(got["Struc"]):
        raise IOError("Sequence and structure of unequal length.")
    if not got["Query"]:
        raise IOError("Input does not contain input sequence and structure.")
    return [got["Query"], got["Struc"]]


def read_molecule_from_file(filename, type):
    try:
        with open(filename, "r") as fh:
            return read_molecule(fh.read(), type)
    except FileNotFoundError as e:
        print("Input file not found.")
        print(e)
        sys.exit(-1)
    except IOError as e:
        print(f"Cannot read input file {filename}.")
        print(e)
        sys.exit(-1)


def breaklines(alilines, width):
    """Split (name, row) pairs into blocks of `width` columns (nonpyx:96-112)."""
    length = len(alilines[0][1])
    return [[(name, row[off:off + width]) for name, row in alilines] for off in range(0, length, width)]


def plot_alignment(*args, **kwargs):
    """SVG rendering (nonpyx:98-367) is visualisation only and outside this engine's scope."""
    raise NotImplementedError("plot_alignment is not part of bialign_b200 (matplotlib figure code is out of scope)")
