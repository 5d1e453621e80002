#!/usr/bin/env python3
"""Command line front-end with the reference's interface (src/bialign.py of s-will/BiAlign): same
arguments (argparse prefix abbreviations included), same `Input:` block, `SCORE:` line and output modes --
the alignment itself is computed on the GPU by bialign_b200."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bialign_b200 import bialignment  # noqa: E402

VERSION_STRING = f"BiAlign {bialignment.__version__}"


def bialign(seqA, seqB, strA, strB, verbose, **args):
    ba = bialignment.BiAligner(seqA, seqB, strA, strB, **args)
    yield "SCORE: " + str(ba.optimize())
    yield ""
    yield from ba.decode_trace()
    if verbose:
        yield from ba.eval_trace()


def build_parser():
    p = argparse.ArgumentParser(description="Bialignment.")
    p.add_argument("seqA", help="sequence A")
    p.add_argument("seqB", help="sequence B")
    p.add_argument("--strA", default=None, help="structure A")
    p.add_argument("--strB", default=None, help="structure B")
    p.add_argument("--nameA", default="A", help="name A")
    p.add_argument("--nameB", default="B", help="name B")
    p.add_argument("-v", "--verbose", action="store_true", help="Verbose")
    p.add_argument("--type", default="RNA", type=str, help="Type of molecule: RNA or Protein")
    p.add_argument("--nodescription", action="store_true",
                   help="Don't prefix the strings in output alignment with descriptions")
    p.add_argument("--outmode", default="default", help="Output mode [call --outmode help for a list of options]")
    p.add_argument("--sequence_match_similarity", type=int, default=100, help="Similarity of matching nucleotides")
    p.add_argument("--sequence_mismatch_similarity", type=int, default=0, help="Similarity of mismatching nucleotides")
    p.add_argument("--structure_weight", type=int, default=400, help="Weighting factor for structure similarity")
    p.add_argument("--gap_opening_cost", type=int, default=0,
                   help="Similarity of opening a gap (turns on affine gap cost if not 0)")
    p.add_argument("--gap_cost", type=int, default=-200, help="Similarity of a single gap position")
    p.add_argument("--shift_cost", type=int, default=-250,
                   help="Similarity of shifting the two scores against each other")
    p.add_argument("--max_shift", type=int, default=2,
                   help="Maximal number of shifts away from the diagonal in either direction")
    p.add_argument("--fileinput", action="store_true", help="Read sequence and structure input from file")
    p.add_argument("--version", action="version", version=VERSION_STRING)
    p.add_argument("--simmatrix", type=str, default=None, help="Similarity matrix")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.fileinput:
        args.seqA, args.strA = bialignment.read_molecule_from_file(args.seqA, args.type)
        args.seqB, args.strB = bialignment.read_molecule_from_file(args.seqB, args.type)
    descr = ["Input:", "seqA\t " + args.seqA, "seqB\t " + args.seqB]
    if args.strA is not None:
        descr.append("strA\t " + args.strA)
    if args.strB is not None:
        descr.append("strB\t " + args.strB)
    print("\n".join(descr))
    if args.outmode == "help":
        print()
        print("Available modes: " + ", ".join(bialignment.BiAligner.outmodes.keys()))
        print()
        sys.exit()
    for line in bialign(**vars(args)):
        print(line)


if __name__ == "__main__":
    main()
