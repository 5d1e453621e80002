"""Deterministic stand-in for the ViennaRNA Python module (test infrastructure, NOT a folding algorithm).

The reference predicts base-pair probabilities with ViennaRNA when no structure is supplied (pyx:345-353:
`RNA.fold_compound(seq)`, `.mfe()`, `.pf()`, `.bpp()`); ViennaRNA is not installed in this image.  This module offers
the same four calls and returns a reproducible pseudo-random probability matrix derived from the sequence, so that the
UNMODIFIED reference can run its probabilistic scoring path here (tests/golden/make_rna_prob_golden.py) and the drop-in
can be checked against it on exactly the same inputs.  The numbers mean nothing biologically."""
import hashlib

import numpy as np


class fold_compound:
    def __init__(self, sequence):
        self.sequence = str(sequence)
        n = len(self.sequence)
        seed = int.from_bytes(hashlib.sha256(self.sequence.encode()).digest()[:8], "little")
        rng = np.random.default_rng(seed)
        p = np.zeros((n + 1, n + 1), dtype=float)
        # a few candidate partners per position, at least three apart, canonical-looking pairs preferred
        good = {("A", "U"), ("U", "A"), ("G", "C"), ("C", "G"), ("G", "U"), ("U", "G")}
        for i in range(1, n + 1):
            for j in range(i + 4, n + 1):
                if rng.random() < 0.12:
                    w = rng.random() * (1.0 if (self.sequence[i - 1], self.sequence[j - 1]) in good else 0.25)
                    p[i, j] = w
        # scale so that every position pairs with total probability <= 0.9
        tot = p.sum(axis=0) + p.sum(axis=1)
        scale = 0.9 / max(0.9, tot.max())
        p *= scale
        self._bpp = p
        # a dot-bracket-like string: the most probable non-crossing greedy pairs (only used for display rows)
        st = ["."] * n
        used = set()
        pairs = sorted(((p[i, j], i, j) for i in range(1, n + 1) for j in range(i + 1, n + 1) if p[i, j] > 0.3), reverse=True)
        chosen = []
        for _, i, j in pairs:
            if i in used or j in used or any((a < i < b < j) or (i < a < j < b) for a, b in chosen):
                continue
            chosen.append((i, j))
            used.update((i, j))
            st[i - 1], st[j - 1] = "(", ")"
        self._structure = "".join(st)

    def mfe(self):
        return (self._structure, -1.0)

    def pf(self):
        return (self._structure.replace("(", "{").replace(")", "}") if False else self._structure, -1.5)

    def bpp(self):
        return [list(map(float, row)) for row in self._bpp]
