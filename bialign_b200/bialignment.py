"""Drop-in for the reference's `bialignment` module (s-will/BiAlign, src/bialignment.pyx).

Same Python surface -- `BiAligner(seqA, seqB, strA, strB, **params)`, `optimize()`,
`traceback()`, `decode_trace()`, `decode_trace_full()`, `eval_trace()`, `outmodes`, the module
helpers -- but the dynamic program (pyx:443-586) runs on the GPU through the C ABI of
include/bialign_b200.h.  There is no CPU implementation of the DP in this package: without the
built CUDA library and an sm_100 device, `optimize()` raises.

Observable behaviour kept from the reference (file:line into src/bialignment.pyx):
  * constructor errors: `ERROR: ...` on stdout + sys.exit(-1) (pyx:207-210, 355-358); KeyError for
    a missing required parameter (pyx:186-190);
  * optimize() returns numpy.int64 (pyx:509 / pyx:471); KeyError for a residue that the similarity
    matrix does not know (pyx:407);
  * traceback() returns forward-ordered columns, lists for the affine model (pyx:568,586), tuples
    for the non-affine one (pyx:526,531), and prints the incomplete-traceback warning (pyx:584-585).
"""
import sys
from math import sqrt

import numpy as np

from . import encoding
from ._capi import BialignError, get_engine
from .encoding import read_simmatrix  # noqa: F401  (re-exported like the reference, pyx:3-4)
from .presentation import (breaklines, consensus_sbpp, consensus_sequence, highlight_sequence_identity,  # noqa: F401
                           mea, parse_dotbracket, plot_alignment, read_molecule, read_molecule_from_file)

__version__ = "0.3"  # surface version of the reference this module mirrors (nonpyx:3)


def guard_case(o, x, max_shift):
    """pyx:133-148: predecessor x - o is non-negative and inside the shift band."""
    return (x[0] - o[0] >= 0 and x[1] - o[1] >= 0 and x[2] - o[2] >= 0 and x[3] - o[3] >= 0
            and abs(x[2] - o[2] - (x[0] - o[0])) <= max_shift and abs(x[3] - o[3] - (x[1] - o[1])) <= max_shift)


def argmin(xs):
    return min(enumerate(xs), key=lambda x: x[1])[0]


def _column_score_affine(state, x, mu1, mu2, beta, gamma, Delta):
    """Score of one column given the source state (behaviour of pyx:84-131); used by eval_trace only."""
    score = Delta * (abs(x[0] - x[2]) + abs(x[1] - x[3]))
    for a, b, mu in ((0, 1, mu1), (2, 3, mu2)):
        if x[a] and x[b]:
            score += mu
        elif x[a] or x[b]:
            score += gamma
            if not (state[a] == x[a] and state[b] == x[b]):
                score += beta
    return score


def _symmetric_pair_matrix(bpp):
    """1-based symmetric pair-probability matrix from ViennaRNA's upper-triangular `bpp`, diagonal = probability of being
    unpaired (behaviour of pyx:327-338; row sums accumulated left to right like the reference's Python sums)."""
    n = len(bpp) - 1
    upper = np.triu(np.asarray(bpp, dtype=float)[: n + 1, : n + 1], k=1)
    upper[0, :] = 0.0
    sb = upper + upper.T
    for i in range(1, n + 1):
        sb[i, i] = 1.0 - (np.cumsum(sb[i, 1:])[-1] if n else 0.0)
    return sb


def _pairing_profile(mol):
    """(up, down, unp) lists over positions 0..n (pyx:366-374): probability of pairing upstream (partner at most i-2),
    downstream, or not at all; sums run left to right like the reference's."""
    n = mol["len"]
    if "sbpp" not in mol:  # supplied dot-bracket string: one-hot profile from the structure classes
        cls = np.concatenate([[encoding.UNP], mol["cls"]])
        return ([float(c == encoding.UP) for c in cls], [float(c == encoding.DOWN) for c in cls],
                [float(c == encoding.UNP) for c in cls])
    sb = mol["sbpp"]
    up = [float(np.cumsum(sb[i, 1:i - 1])[-1]) if i - 1 > 1 else 0.0 for i in range(n + 1)]
    down = [float(np.cumsum(sb[i, i + 1:n + 1])[-1]) if i + 1 <= n else 0.0 for i in range(n + 1)]
    unp = [1.0 - up[i] - down[i] for i in range(n + 1)]
    return up, down, unp


class BiAligner:
    nl = 14
    outmodes = {
        "default": [1, 3, 6, 8, 12, 13],
        "sorted": [0, 1, 5, 3, 2, 4, nl] + [7, 6, 10, 8, 9, 11, nl] + [12, 13],
        "sorted_sym": [0, 1, 3, 2, 5, 4, nl] + [6, 7, 9, 8, 11, 10, nl] + [12, 13],
        "sorted_terse": [1, 5, 3, 4, nl] + [6, 10, 8, 11, nl] + [12, 13],
        "raw": [1, 3, 7, 9],
        "raw_struct": list(range(4)) + list(range(6, 10)),
        "full": range(nl),
    }

    def __init__(self, seqA, seqB, strA, strB, **params):
        # parameter capture of pyx:179-193: molecule A before molecule B, then the three costs, then the matrix -- a missing
        # key raises KeyError at the same point as in the reference
        self._params = params
        self.molA, self.molB = [self._preprocess_seq(q, st) for q, st in ((seqA, strA), (seqB, strB))]
        self.gamma, self.beta, self.max_shift = [params[key] for key in ("gap_cost", "gap_opening_cost", "max_shift")]
        self._simmatrix = read_simmatrix(params["simmatrix"]) if params["simmatrix"] else None
        self._score = self._trace = None
        self._complete = True

    # ------------------------------------------------------------------ helpers of the surface
    @property
    def _is_rna(self):
        return self._params["type"] == "RNA"

    @property
    def _affine(self):
        return self.beta != 0

    @staticmethod
    def error(text):
        print("ERROR:", text)
        sys.exit(-1)

    def _preprocess_seq(self, sequence, structure):
        """The molecule record the rest of the class works on (pyx:340-376): sequence, length, structure and -- for RNAs --
        the pairing profile of every position, from the supplied dot-bracket string or from ViennaRNA."""
        mol = {"seq": str(sequence)}
        mol["len"] = len(mol["seq"])
        if structure is not None:
            if len(structure) != len(sequence):
                self.error("Provided structure and sequence must have the same length.")
            mol["structure"] = structure
            if self._is_rna:
                mol["partner"] = encoding.dotbracket_partners(structure)
                mol["cls"] = encoding.rna_structure_classes(structure)
        elif not self._is_rna:
            self.error("Structures have to be provided when aligning proteins")
        else:
            mol.update(self._predicted_structure(mol["seq"]))
        if self._is_rna:
            mol["up"], mol["down"], mol["unp"] = _pairing_profile(mol)
        return mol

    @staticmethod
    def _predicted_structure(sequence):
        """No structure supplied for an RNA: base-pair probabilities from ViennaRNA, the calls of pyx:345-353 in their order
        (ModuleNotFoundError when the module is absent, like the reference).  The probabilistic structure similarity is
        evaluated on the host with the reference's floating-point formula and handed to the engine as an integer matrix."""
        import RNA

        fc = RNA.fold_compound(sequence)
        mfe = fc.mfe()
        pf = fc.pf()
        sbpp = _symmetric_pair_matrix(fc.bpp())
        return {"mfe": mfe, "pf": pf, "sbpp": sbpp, "mea": mea(sbpp), "structure": pf[0], "predicted": True}

    # scoring functions, 1-based like the reference (pyx:405-440); used by eval_trace
    def mu1(self, i, j):
        a, b = self.molA["seq"][i - 1], self.molB["seq"][j - 1]
        if self._simmatrix:
            return self._simmatrix[a][b]
        if a == b:
            return self._params["sequence_match_similarity"]
        return self._params["sequence_mismatch_similarity"]

    def mu2(self, i, j):
        if self._is_rna:
            # pyx:416-423; with supplied structures the profiles are 0/1 and this is w * [class_A(i) == class_B(j)]
            A, B = self.molA, self.molB
            return int(self._params["structure_weight"] * (sqrt(A["up"][i] * B["up"][j]) + sqrt(A["down"][i] * B["down"][j]) +
                                                           sqrt(A["unp"][i] * B["unp"][j])))
        if self.molA["structure"][i - 1] == self.molB["structure"][j - 1]:
            return self._params["structure_weight"]
        return 0

    def _mu2_matrix(self):
        """int32 matrix mu2(k, l), k = 1..n, l = 1..m, when a molecule has a predicted (probabilistic) profile, else None."""
        if not (self._is_rna and (self.molA.get("predicted") or self.molB.get("predicted"))):
            return None
        A, B = self.molA, self.molB
        total = 0.0
        for key in ("up", "down", "unp"):
            prod = np.outer(np.asarray(A[key][1:], dtype=float), np.asarray(B[key][1:], dtype=float))
            if (prod < 0).any():
                raise ValueError("math domain error")  # math.sqrt of a negative product (pyx:419-421)
            total = total + np.sqrt(prod)
        return (self._params["structure_weight"] * total).astype(np.int64).astype(np.int32)

    # ------------------------------------------------------------------ the hot path (GPU)
    def _encoded(self):
        """(residues A, residues B, classes A, classes B, similarity table) for the C ABI."""
        sa, sb = self.molA["seq"], self.molB["seq"]
        if self._simmatrix:
            symbols, table, known = encoding.simmatrix_table(self._simmatrix)
            # the reference fails lazily with KeyError at the first unknown residue pair it scores
            # (pyx:407; the first pair evaluated is (seqA[-1], seqB[-1]))
            lut = {c: i for i, c in enumerate(symbols)}
            order_a = ([sa[-1]] if sa else []) + list(sa)
            order_b = ([sb[-1]] if sb else []) + list(sb)
            for ch in order_a:
                if ch not in self._simmatrix:
                    raise KeyError(ch)
            cols_ok = set(c for c in symbols if all(known[lut[a], lut[c]] for a in set(sa)))
            for ch in order_b:
                if ch not in cols_ok:
                    raise KeyError(ch)
            ra = encoding.encode_residues(sa, symbols)
            rb = encoding.encode_residues(sb, symbols)
        else:
            ra, rb = encoding.encode_bytes(sa), encoding.encode_bytes(sb)
            used, inv = np.unique(np.concatenate([ra, rb]), return_inverse=True)
            ra, rb = inv[:len(ra)].astype(np.uint8), inv[len(ra):].astype(np.uint8)
            table = encoding.match_table(self._params["sequence_match_similarity"],
                                         self._params["sequence_mismatch_similarity"], nsym=max(len(used), 1))
        if self._is_rna:
            # (a molecule with a predicted profile has no classes: its mu2 comes from the uploaded matrix)
            ca = self.molA.get("cls", np.zeros(self.molA["len"], dtype=np.uint8))
            cb = self.molB.get("cls", np.zeros(self.molB["len"], dtype=np.uint8))
        else:
            ca, cb = encoding.encode_bytes(self.molA["structure"]), encoding.encode_bytes(self.molB["structure"])
        return ra, rb, ca, cb, table

    def optimize(self):
        n, m = self.molA["len"], self.molB["len"]
        if (n == 0) != (m == 0):
            raise IndexError("string index out of range")  # the reference's seq[-1] on an empty string (pyx:407)
        ra, rb, ca, cb, table = self._encoded()
        eng = get_engine()
        eng.apply_options()  # the engine is shared per device: run with automatic settings whatever was set before
        eng.set_scoring(table, self._params["structure_weight"], self.beta, self.gamma, self._params["shift_cost"],
                        self.max_shift)
        res = np.concatenate([ra, rb]).astype(np.uint8)
        cls = np.concatenate([ca, cb]).astype(np.uint8)
        off = np.array([0, n, n + m], dtype=np.int64)
        eng.load_sequences(res, cls, off)
        eng.load_pairs(np.array([0], dtype=np.int32), np.array([1], dtype=np.int32))
        mu2 = self._mu2_matrix()
        if mu2 is not None:
            eng.set_pair_mu2([mu2])
        eng.run(want_trace=True)
        self._score = np.int64(eng.fetch_scores()[0])
        cols, offsets, complete = eng.fetch_traces()
        self._trace = cols[offsets[0]:offsets[1]].copy()
        self._complete = bool(complete[0])
        return self._score

    def traceback(self):
        if self._trace is None:
            raise TypeError("'NoneType' object is not subscriptable")  # reference: traceback() before optimize()
        if self.molA["len"] == 0 and self.molB["len"] == 0:
            raise IndexError("string index out of range")  # pyx:555 -> pyx:260 -> pyx:407 on empty strings
        cols = [[(c >> 3) & 1, (c >> 2) & 1, (c >> 1) & 1, c & 1] for c in self._trace.tolist()]
        if self._affine:
            if not self._complete:
                print("WARNING: incomplete traceback. Alignment could be garbage.")
            return cols
        return [tuple(c) for c in cols]

    # ------------------------------------------------------------------ presentation (host only)
    @staticmethod
    def _transfer_gaps(alistr, seqstr):
        """`seqstr` spread over the non-gap columns of `alistr` (its gaps kept)."""
        symbols = iter(seqstr)
        return "".join("-" if c == "-" else next(symbols) for c in alistr)

    @staticmethod
    def auto_complete(x, xs):
        """First of `xs` (alphabetically) that starts with `x`; `x` itself when none does."""
        return next((y for y in sorted(xs) if y.startswith(x)), x)

    def _sbpp(self, mol):
        """Symmetric pair matrix with unpaired probability on the diagonal for a fixed structure
        (what pyx:378-392 builds); only needed for the RNA consensus-structure rows."""
        if "sbpp" in mol:
            return mol["sbpp"]
        n = mol["len"]
        m = np.zeros((n + 1, n + 1), dtype=float)
        partner = mol["partner"]
        for i in range(1, n + 1):
            p = int(partner[i])
            ch = mol["structure"][i - 1]
            if p:
                m[i, p] = 1.0
            elif ch not in "()":
                m[i, i] = 1.0
        return m

    def _consensus_structure(self, ssA, ssB):
        """Consensus-structure row of one of the two alignments, from the gapped structure strings of A and B."""
        if not self._is_rna:
            return consensus_sequence(ssA, ssB)
        pairs = consensus_sbpp(alistrA=ssA, alistrB=ssB, sbppA=self._sbpp(self.molA), sbppB=self._sbpp(self.molB))
        return mea(pairs, brackets="[]")[0]

    def decode_trace_full(self, trace=None):
        """Trace -> 14 (name, row) pairs, the row set of pyx:633-707: for the sequence alignment (columns x0, x1) and then
        for the structure alignment (columns x2, x3) the six rows `A ss`, `A`, `B ss`, `B`, `consensus ss`, `consensus`,
        followed by the two shift rows."""
        if trace is None:
            trace = self.traceback()
        cols = np.asarray(trace, dtype=np.int64).reshape(-1, 4)
        mols = (self.molA, self.molB)
        nameA, nameB = self._params["nameA"], self._params["nameB"]
        rows, gapped = [], []
        for copy in (0, 1):
            block = []
            for which, mol in enumerate(mols):
                advances = cols[:, 2 * copy + which] == 1
                block.append("".join(np.where(advances, np.array(list(mol["seq"]) + ["-"], dtype="<U1")[
                    np.minimum(np.cumsum(advances) - 1, mol["len"])], "-")) if len(cols) else "")
            gapped.append(block)
            ss = [self._transfer_gaps(row, mol["structure"]) for row, mol in zip(block, mols)]
            rows += [(nameA + " ss", ss[0]), (nameA, block[0]), (nameB + " ss", ss[1]), (nameB, block[1]),
                     ("consensus ss", self._consensus_structure(ss[0], ss[1])),
                     ("consensus", consensus_sequence(block[0], block[1]))]
        for which, name in enumerate((nameA, nameB)):
            first, second = gapped[0][which], gapped[1][which]
            rows.append((name + " shifts", "".join("." if (a == "-") == (b == "-") else (">" if a == "-" else "<")
                                                   for a, b in zip(first, second))))
        return rows

    def decode_trace(self, trace=None):
        """The rows selected by the output mode (pyx:168-177, 709-743), each prefixed with its left-justified name unless
        `nodescription`; index 14 of an output mode is the empty separator line."""
        rows = self.decode_trace_full(trace)
        if self._params.get("nodescription"):
            lines = [text for _, text in rows]
        else:
            pad = 4 + max(len(label) for label, _ in rows)
            lines = [label.ljust(pad) + text for label, text in rows]
        lines.append("")
        requested = self._params.setdefault("outmode", "default")
        layout = self.outmodes.get(self.auto_complete(requested, self.outmodes.keys()))
        if layout is None:
            print("WARNING: unknown output mode. Expect one of " + str(list(self.outmodes.keys())))
            layout = self.outmodes["sorted"]
        return [lines[i] for i in layout]

    # ------------------------------------------------------------------ trace evaluation (-v)
    def _nonaffine_cases(self, idx):
        i, j, k, l = idx
        m1, m2 = self.mu1(i, j), self.mu2(k, l)
        g, D = self.gamma, self._params["shift_cost"]
        return [((1, 1, 1, 1), m1 + m2), ((1, 0, 1, 0), g + g), ((0, 1, 0, 1), g + g), ((1, 1, 0, 0), m1 + D),
                ((0, 0, 1, 1), m2 + D), ((1, 0, 0, 0), g + D), ((0, 1, 0, 0), g + D), ((0, 0, 1, 0), g + D),
                ((0, 0, 0, 1), g + D), ((1, 0, 1, 1), g + m2 + D), ((0, 1, 1, 1), g + m2 + D),
                ((1, 1, 1, 0), g + m1 + D), ((1, 1, 0, 1), g + m1 + D)]

    def eval_trace(self, trace=None):
        """Per-column score listing (pyx:745-832).  For the non-affine model the reference prints the
        DP value of the previous cell plus the column score; along an optimal trace that is the
        running total, which is what is printed here."""
        if trace is None:
            trace = self.traceback()
        Delta = self._params["shift_cost"]
        idx = [0, 0, 0, 0]
        total = 0
        if self._affine:
            state = [1, 1, 1, 1]
            for y in trace:
                y = list(y)
                for q in range(4):
                    idx[q] += y[q]
                i, j, k, l = idx
                score = _column_score_affine(state, y, self.mu1(i, j), self.mu2(k, l), self.beta, self.gamma, Delta)
                total += score
                if y[0] or y[1]:
                    state[0], state[1] = y[0], y[1]
                if y[2] or y[3]:
                    state[2], state[3] = y[2], y[3]
                yield " ".join(str(item) for item in [idx, y, score, "-->", total])
            return
        for y in trace:
            for q in range(4):
                idx[q] += y[q]
            for x, sc in self._nonaffine_cases(idx):
                if x == y:
                    total += sc
                    yield " ".join(str(item) for item in [idx, y, sc, "-->", total])
                    break
