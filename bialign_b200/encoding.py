"""Host-side encoders: strings -> the uint8 residue / structure-class arrays and int32 similarity
table of include/bialign_b200.h.  Behaviour follows the reference's scoring inputs
(src/bialignment.pyx:340-440, src/bialignment_nonpyx.py:5-58)."""
import numpy as np

# BLOSUM62 over the reference's 24-letter alphabet (nonpyx:5-30 / Data/BLOSUM62.txt hold the same
# numbers); row/column order = ALPHABET.  Entries are multiplied by `scale` (100) on load.
ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"
_B62 = (
    "4,-1,-2,-2,0,-1,-1,0,-2,-1,-1,-1,-1,-2,-1,1,0,-3,-2,0,-2,-1,0,-4;"
    "-1,5,0,-2,-3,1,0,-2,0,-3,-2,2,-1,-3,-2,-1,-1,-3,-2,-3,-1,0,-1,-4;"
    "-2,0,6,1,-3,0,0,0,1,-3,-3,0,-2,-3,-2,1,0,-4,-2,-3,3,0,-1,-4;"
    "-2,-2,1,6,-3,0,2,-1,-1,-3,-4,-1,-3,-3,-1,0,-1,-4,-3,-3,4,1,-1,-4;"
    "0,-3,-3,-3,9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1,-3,-3,-2,-4;"
    "-1,1,0,0,-3,5,2,-2,0,-3,-2,1,0,-3,-1,0,-1,-2,-1,-2,0,3,-1,-4;"
    "-1,0,0,2,-4,2,5,-2,0,-3,-3,1,-2,-3,-1,0,-1,-3,-2,-2,1,4,-1,-4;"
    "0,-2,0,-1,-3,-2,-2,6,-2,-4,-4,-2,-3,-3,-2,0,-2,-2,-3,-3,-1,-2,-1,-4;"
    "-2,0,1,-1,-3,0,0,-2,8,-3,-3,-1,-2,-1,-2,-1,-2,-2,2,-3,0,0,-1,-4;"
    "-1,-3,-3,-3,-1,-3,-3,-4,-3,4,2,-3,1,0,-3,-2,-1,-3,-1,3,-3,-3,-1,-4;"
    "-1,-2,-3,-4,-1,-2,-3,-4,-3,2,4,-2,2,0,-3,-2,-1,-2,-1,1,-4,-3,-1,-4;"
    "-1,2,0,-1,-3,1,1,-2,-1,-3,-2,5,-1,-3,-1,0,-1,-3,-2,-2,0,1,-1,-4;"
    "-1,-1,-2,-3,-1,0,-2,-3,-2,1,2,-1,5,0,-2,-1,-1,-1,-1,1,-3,-1,-1,-4;"
    "-2,-3,-3,-3,-2,-3,-3,-3,-1,0,0,-3,0,6,-4,-2,-2,1,3,-1,-3,-3,-1,-4;"
    "-1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4,7,-1,-1,-4,-3,-2,-2,-1,-2,-4;"
    "1,-1,1,0,-1,0,0,0,-1,-2,-2,0,-1,-2,-1,4,1,-3,-2,-2,0,0,0,-4;"
    "0,-1,0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1,1,5,-2,-2,0,-1,-1,0,-4;"
    "-3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1,1,-4,-3,-2,11,2,-3,-4,-3,-2,-4;"
    "-2,-2,-2,-3,-2,-1,-2,-3,2,-1,-1,-2,-1,3,-3,-2,-2,2,7,-1,-3,-2,-1,-4;"
    "0,-3,-3,-3,-1,-2,-2,-3,-3,3,1,-2,1,-1,-2,-2,0,-3,-1,4,-3,-2,-1,-4;"
    "-2,-1,3,4,-3,0,1,-1,0,-3,-4,0,-3,-3,-2,0,-1,-4,-3,-3,4,1,-1,-4;"
    "-1,0,0,1,-3,3,4,-2,0,-3,-3,1,-1,-3,-1,0,-1,-3,-2,-2,1,4,-1,-4;"
    "0,-1,-1,-1,-2,-1,-1,-1,-1,-1,-1,-1,-1,-1,-2,0,0,-2,-1,-1,-1,-1,-1,-4;"
    "-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,1"
)


def blosum62_rows():
    return [[int(v) for v in row.split(",")] for row in _B62.split(";")]


def read_simmatrix(filename, scale=100):
    """dict-of-dicts similarity matrix, entries * scale (interface of nonpyx:33-58).

    "BLOSUM62" selects the built-in table; otherwise `filename` is a text file whose header row
    starts with '-' followed by the column symbols, then one row per symbol."""
    if filename == "BLOSUM62":
        rows = blosum62_rows()
        return {a: {b: scale * rows[i][j] for j, b in enumerate(ALPHABET)} for i, a in enumerate(ALPHABET)}
    with open(filename, "r") as fh:
        return _parse_matrix_text(fh, scale)


def _parse_matrix_text(lines, scale):
    """Single pass over the text of a matrix file.  The header (first field '-') names the columns and fixes how many
    lines can belong to the table: with c columns nothing beyond line index c is looked at.  Every other line is a row:
    its symbol, then the values of the columns in header order (a short row simply defines fewer entries)."""
    columns, row_order, table, last = None, [], {}, None
    for index, text in enumerate(lines):
        if last is not None and index > last:
            break
        symbol, *values = text.split()
        if symbol == "-":
            columns, last = values, (len(values) or None)
            continue
        row_order.append(symbol)
        table[symbol] = dict(zip(columns, (scale * int(v) for v in values)))
    if columns != row_order:
        print("ERROR while reading simmatrix {filename}.")  # the reference prints exactly this (no interpolation)
    return table


def simmatrix_table(matrix):
    """dict-of-dicts -> (symbols str, dense int32 table) with symbols in insertion order."""
    syms = list(matrix.keys())
    cols = []
    for r in syms:
        for c in matrix[r]:
            if c not in syms and c not in cols:
                cols.append(c)
    allsyms = syms + cols
    t = np.zeros((len(allsyms), len(allsyms)), dtype=np.int32)
    known = np.zeros((len(allsyms), len(allsyms)), dtype=bool)
    for i, r in enumerate(allsyms):
        for j, c in enumerate(allsyms):
            if r in matrix and c in matrix[r]:
                t[i, j] = matrix[r][c]
                known[i, j] = True
    return allsyms, t, known


def match_table(match, mismatch, nsym=256):
    t = np.full((nsym, nsym), int(mismatch), dtype=np.int32)
    idx = np.arange(nsym)
    t[idx, idx] = int(match)
    return t


def encode_bytes(text):
    """Raw one-byte-per-character codes (for match/mismatch scoring and protein structure symbols)."""
    try:
        return np.frombuffer(text.encode("latin-1"), dtype=np.uint8).copy()
    except UnicodeEncodeError:
        # characters beyond latin-1: map each distinct character to a small code, equality-preserving
        table = {}
        return np.array([table.setdefault(ch, len(table) % 256) for ch in text], dtype=np.uint8)


def encode_residues(seq, symbols):
    """Codes of `seq` over `symbols` (list of one-character strings); KeyError on an unknown residue
    like the reference's dict lookup (pyx:407)."""
    lut = {c: i for i, c in enumerate(symbols)}
    return np.array([lut[c] for c in seq], dtype=np.uint8)


def dotbracket_partners(structure):
    """Partner (1-based, 0 = none) of every position of a dot-bracket string, as pyx:378-392 builds
    the base-pair matrix: '(' opens, ')' closes the innermost open bracket (IndexError if none),
    every other character and every unclosed '(' is unpaired."""
    n = len(structure)
    partner = np.zeros(n + 1, dtype=np.int64)
    stack = []
    for i, ch in enumerate(structure):
        if ch == "(":
            stack.append(i)
        elif ch == ")":
            j = stack.pop()
            partner[i + 1] = j + 1
            partner[j + 1] = i + 1
    return partner


UNP, UP, DOWN = 0, 1, 2


def rna_structure_classes(structure):
    """Per-position class of a supplied dot-bracket structure: with 0/1 pair probabilities the
    reference's up/down/unp profile (pyx:366-374) is one-hot, so mu2 (pyx:416-423) reduces to
    w * [class_A(k) == class_B(l)].  `up` needs the partner at least two positions to the left
    (pyx:368 sums j in range(1, i-1)); `down` a partner to the right; everything else is unpaired."""
    partner = dotbracket_partners(structure)
    pos = np.arange(len(structure) + 1)
    cls = np.full(len(structure) + 1, UNP, dtype=np.uint8)
    cls[(partner > 0) & (partner <= pos - 2)] = UP
    cls[(partner > 0) & (partner >= pos + 1)] = DOWN
    return cls[1:].copy()
