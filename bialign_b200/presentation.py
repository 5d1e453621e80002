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
    """0-based partner of every position, -1 for unpaired (behaviour of pyx:911-922; IndexError on an unbalanced ')')."""
    from .encoding import dotbracket_partners

    return [int(p) - 1 for p in dotbracket_partners(dbstr)[1:]]


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
    """Maximum-expected-accuracy structure of a symmetric pair matrix with unpaired weights on the diagonal; returns
    (structure string, score).  Spec (pyx:836-886): F[i][j] = best over "j unpaired" (weight sbpp[j][j]) and "j pairs
    with k" (k < j - 3, weight 2*gamma*sbpp[k][j]) of F[i][k-1] + F[k+1][j-1] + weight, where only pairs (k, j) that
    strictly beat every alternative for the interval [k, j] are ever considered, ties go to "j unpaired" and then to the
    largest k, and an interval whose best value is 0 stays open.  Implementation: per end position j the admissible
    splits are kept as two growing arrays (k and the value of closing at k), so one interval is a single vectorised
    max over them; the structure is then read off the argmax table interval by interval."""
    sbpp = np.asarray(sbpp, dtype=float)
    n = len(sbpp) - 1
    F = np.zeros((n + 2, n + 2), dtype=float)
    arg = np.zeros((n + 2, n + 2), dtype=np.int64)      # 0 = nothing gained on this interval
    split_at = [np.empty(0, dtype=np.int64) for _ in range(n + 1)]
    split_val = [np.empty(0, dtype=float) for _ in range(n + 1)]
    for i in range(n, 0, -1):
        split_at[i] = np.append(split_at[i], i)          # "i unpaired" is always admissible for intervals ending at i
        split_val[i] = np.append(split_val[i], sbpp[i, i])
        for j in range(i, n + 1):
            totals = F[i, split_at[j] - 1] + split_val[j]
            w = int(np.argmax(totals))                   # first maximum = order of admission: j itself, then larger k first
            if totals[w] > 0.0:
                F[i, j], arg[i, j] = totals[w], split_at[j][w]
            if j - i > 3:
                closed = F[i + 1, j - 1] + 2 * gamma * sbpp[i, j]
                if closed > F[i, j]:                     # (i, j) strictly better than anything else: admit it for rows above
                    split_at[j] = np.append(split_at[j], i)
                    split_val[j] = np.append(split_val[j], closed)
                    F[i, j], arg[i, j] = closed, i
    opening, closing = brackets[0], brackets[1]
    out = ["."] * (n + 1)
    todo = [(1, n)] if n >= 1 else []
    while todo:
        i, j = todo.pop()
        while j - i > 3 and arg[i, j] == j:               # trailing unpaired positions
            j -= 1
        k = int(arg[i, j]) if j - i > 3 else 0
        if k == 0:
            continue
        out[k], out[j] = opening, closing
        todo.append((k + 1, j - 1))
        if k > i:
            todo.append((i, k - 1))
    return "".join(out[1:]), (F[1, n] if n >= 1 else 0.0)


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
    """[sequence, structure] of a CFSSP-format file (nonpyx:80-93).  Like the reference's, this never raises for an
    unreadable input: a missing file and any other I/O problem (incl. a molecule type read_molecule has no file format
    for) print their message, then the exception text, and end the process with status -1."""
    try:
        with open(filename, "r") as fh:
            text = fh.read()
        return read_molecule(text, type)
    except OSError as e:  # IOError is OSError; FileNotFoundError is the one case with a message of its own
        print("Input file not found." if isinstance(e, FileNotFoundError) else f"Cannot read input file {filename}.")
        print(e)
        sys.exit(-1)


def breaklines(alilines, width):
    """Split (name, row) pairs into blocks of `width` columns (nonpyx:96-112)."""
    length = len(alilines[0][1])
    return [[(name, row[off:off + width]) for name, row in alilines] for off in range(0, length, width)]


def plot_alignment(*args, **kwargs):
    """SVG rendering (nonpyx:98-367) is visualisation only and outside this engine's scope."""
    raise NotImplementedError("plot_alignment is not part of bialign_b200 (matplotlib figure code is out of scope)")
