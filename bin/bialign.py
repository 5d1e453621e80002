#!/usr/bin/env python3
"""Command line front-end with the interface of the reference's `bialign.py`: same positional and optional
arguments (argparse prefix abbreviations work: `--structure 400`, `--filein`), same `Input:` block, `SCORE:`
line and output modes -- the alignment itself is computed on the GPU by bialign_b200.  Checked against
transcripts of the reference CLI (tests/golden/cli_outputs.json)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bialign_b200 import bialignment  # noqa: E402

# (flag, default, type or None for store_true, help) -- the option surface of the reference CLI
STRING_OPTS = [("--strA", None, "structure A"), ("--strB", None, "structure B"), ("--nameA", "A", "name A"),
               ("--nameB", "B", "name B"), ("--type", "RNA", "Type of molecule: RNA or Protein"),
               ("--outmode", "default", "Output mode [call --outmode help for a list of options]"),
               ("--simmatrix", None, "Similarity matrix")]
INT_OPTS = [("--sequence_match_similarity", 100, "Similarity of matching nucleotides"),
            ("--sequence_mismatch_similarity", 0, "Similarity of mismatching nucleotides"),
            ("--structure_weight", 400, "Weighting factor for structure similarity"),
            ("--gap_opening_cost", 0, "Similarity of opening a gap (turns on affine gap cost if not 0)"),
            ("--gap_cost", -200, "Similarity of a single gap position"),
            ("--shift_cost", -250, "Similarity of shifting the two scores against each other"),
            ("--max_shift", 2, "Maximal number of shifts away from the diagonal in either direction")]
FLAG_OPTS = [("--nodescription", "Don't prefix the strings in output alignment with descriptions"),
             ("--fileinput", "Read sequence and structure input from file")]


def build_parser():
    p = argparse.ArgumentParser(description="Bialignment.")
    for pos in ("seqA", "seqB"):
        p.add_argument(pos, help="sequence " + pos[-1])
    for flag, default, text in STRING_OPTS:
        p.add_argument(flag, default=default, type=str, help=text)
    for flag, default, text in INT_OPTS:
        p.add_argument(flag, default=default, type=int, help=text)
    for flag, text in FLAG_OPTS:
        p.add_argument(flag, action="store_true", help=text)
    p.add_argument("-v", "--verbose", action="store_true", help="Verbose")
    p.add_argument("--version", action="version", version=f"BiAlign {bialignment.__version__}")
    return p


def run(opts):
    """Lines the CLI prints after the input block: score, blank line, decoded alignment, optional -v listing."""
    params = dict(opts)
    seqs = [params.pop(k) for k in ("seqA", "seqB", "strA", "strB")]
    verbose = params.pop("verbose")
    aligner = bialignment.BiAligner(*seqs, **params)
    lines = ["SCORE: " + str(aligner.optimize()), ""]
    lines += aligner.decode_trace()
    if verbose:
        lines += list(aligner.eval_trace())
    return lines


def main(argv=None):
    opts = vars(build_parser().parse_args(argv))
    if opts["fileinput"]:
        for mol in "AB":
            opts["seq" + mol], opts["str" + mol] = bialignment.read_molecule_from_file(opts["seq" + mol], opts["type"])
    print("\n".join(["Input:"] + [f"{key}\t {opts[key]}" for key in ("seqA", "seqB", "strA", "strB") if opts[key] is not None]))
    if opts["outmode"] == "help":
        print("\nAvailable modes: " + ", ".join(bialignment.BiAligner.outmodes.keys()) + "\n")
        sys.exit()
    for line in run(opts):
        print(line)


if __name__ == "__main__":
    main()
