#!/usr/bin/env python3
"""Golden CLI transcripts from the UNMODIFIED reference CLI (src/bialign.py run against the compiled
reference in oracle/_ref).  Build container only.  Writes tests/golden/cli_outputs.json."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "..", "oracle", "_ref")
CLI = "/root/reference/src/bialign.py"
PROT = ["RAKLPLKEKKLTATANYHPGIRYIMTGYSAKYIYSSTYARFR", "KAKLPLKEKKLTRTANYHPGIRYIMTGYSAKRIYSSTYAYFR",
        "--strA", "CHHHHHHHHHHHHHCCCCTCEEEEEEECCTCEEEEEEEECCC", "--strB", "HHHHHHHHHHHHCCCCCCTCEEEEEEECCCCCEEEEEEEECC",
        "--type", "Protein", "--shift_cost", "-150", "--structure_weight", "800", "--simmatrix", "BLOSUM62",
        "--gap_opening_cost", "-150", "--gap_cost", "-50", "--max_shift", "1"]
RNA = ["GCGGGGGAUAUCCCCAUCG", "GGGGAUAUCCCCAUCG", "--strA", "...(((.....))).....", "--strB", ".(((.....)))....",
       "--structure", "400", "--gap_opening_cost", "-200", "--gap_cost", "-50", "--max_shift", "1", "--shift_cost", "-150"]
cases = [PROT + ["--outmode", m] for m in ["default", "sorted", "sorted_sym", "sorted_terse", "raw", "raw_struct", "full", "sor"]]
cases += [PROT + ["--outmode", "sorted", "--nodescription"], PROT + ["-v"], RNA, RNA + ["--outmode", "full"], RNA + ["-v"],
          ["GCGGGGGAUAUCCCCAUCG", "GGGGAUAUCCCCAUCG", "--strA", "...(((.....))).....", "--strB", ".(((.....)))....", "-v"],
          ["A", "A", "--outmode", "help"]]
# --fileinput: two small files in the CFSSP report format (Query / Struc rows), committed next to this script; argv holds
# the file names relative to tests/golden (the test and this script both run the CLI from there)
cases += [["small_A.cfssp", "small_B.cfssp", "--filein", "--type", "Protein", "--shift_cost", "-150", "--structure_weight", "800",
           "--simmatrix", "BLOSUM62", "--gap_opening_cost", "-150", "--gap_cost", "-50", "--max_shift", "1", "--outmode", "sorted"]]
out = []
env = dict(os.environ, PYTHONPATH=REF)
for argv in cases:
    r = subprocess.run([sys.executable, CLI] + argv, capture_output=True, text=True, env=env, cwd=HERE)
    out.append({"argv": argv, "stdout": r.stdout, "rc": r.returncode})
    print(argv[-2:], r.returncode, len(r.stdout), file=sys.stderr)
json.dump(out, open(os.path.join(HERE, "cli_outputs.json"), "w"), indent=0)
