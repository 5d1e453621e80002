#!/usr/bin/env python3
"""Registers / stack / static shared memory (cuobjdump -res-usage) and SASS mnemonic counts (cuobjdump -sass) of the
kernels the named shapes run, read from the built in-tree library.  No GPU needed.
    python scripts/sass_evidence.py > profiles/<tag>_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bialign_b200", "libbialign_b200.so")
NAMES = ["S", "TRACE", "PAD", "BNEG", "LONG", "P16", "NA", "CHAIN", "REBASE", "IOW", "TILED"]
# (label, template arguments of fill_systolic_kernel) -- the flavours behind the numbers of DESIGN.md section 7
PICK = [
    ("config 3 / bench kernel: batch, score + traceback", dict(S=2, TRACE=1, BNEG=1)),
    ("config 3 score only", dict(S=2, BNEG=1)),
    ("config 4: 16-bit pair mode", dict(S=2, BNEG=1, P16=1)),
    ("config 1 x 100k: chained short pairs", dict(S=1, TRACE=1, BNEG=1, CHAIN=1)),
    ("config 2: single long pair, I/O warp", dict(S=1, TRACE=1, BNEG=1, LONG=1, IOW=1)),
    ("config 5: single long pair, I/O warp", dict(S=3, TRACE=1, BNEG=1, LONG=1, IOW=1)),
    ("gangs of long pairs (four-warp CTA)", dict(S=2, TRACE=1, BNEG=1, LONG=1)),
    ("rebased wide-range trace launch", dict(S=2, TRACE=1, BNEG=1, REBASE=1)),
]
COUNT = ["VIADDMNMX", "VIMNMX3", "VIMNMX", "IADD3", "IMAD", "LOP3", "SHF", "SEL", "SHFL", "LDS", "STS", "LDGSTS", "LDG", "STG",
         "BAR", "BSSY", "HMMA", "UTCHMMA", "UTCIMMA"]


def run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True, check=True).stdout


def mangled(args):
    a = dict.fromkeys(NAMES, 0)
    a.update(args)
    return "fill_systolic_kernelILi%dE" % a["S"] + "".join("Lb%dE" % a[n] for n in NAMES[1:]) + "EEvNS_7SysArgsE"


def main():
    if not os.path.exists(LIB):
        sys.exit("build the library first: python -m bialign_b200.build")
    res = {}
    cur = None
    for line in run("cuobjdump", "-res-usage", LIB).splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            res[cur] = dict(kv.split(":") for kv in line.split() if ":" in kv and not kv.startswith("CONSTANT"))
            cur = None
    print("# registers / stack / static shared memory and SASS mnemonic counts of the final round-2 build (nvcc 12.9, sm_100a),")
    print("# read from bialign_b200/libbialign_b200.so by scripts/sass_evidence.py; template arguments <%s>" % ", ".join(NAMES))
    print("# %d fill_systolic_kernel instantiations in the library; the ones behind the measured shapes:\n" %
          sum("fill_systolic_kernel" in k for k in res))
    for label, args in PICK:
        tail = mangled(args)
        fn = next((k for k in res if k.endswith(tail)), None)
        if fn is None:
            print("%-55s (not instantiated: %s)" % (label, tail))
            continue
        r = res[fn]
        sass = run("cuobjdump", "-sass", "-fun", fn, LIB)
        ops = collections.Counter()
        n = 0
        for line in sass.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops[m.group(1)] += 1
                n += 1
        counts = " ".join("%s %d" % (k, ops.get(k, 0)) for k in COUNT)
        s16 = sum(1 for line in sass.splitlines() if ".S16x2" in line)
        print("%s\n  <%s>\n  REG %s  STACK %s  static SHARED %s;  %d SASS instructions (whole kernel, all three forms of the iteration)\n  %s%s\n" %
              (label, ", ".join(str(dict(dict.fromkeys(NAMES, 0), **args)[k]) for k in NAMES), r.get("REG"), r.get("STACK"),
               r.get("SHARED"), n, counts, ("  (.S16x2 forms: %d)" % s16) if s16 else ""))
    for k in sorted(res):
        if "fill_systolic_kernel" in k:
            continue
        print("%-80s REG %s STACK %s" % (run("c++filt", k).strip()[:80], res[k].get("REG"), res[k].get("STACK")))
    spill = sorted((int(v.get("STACK", 0)), k) for k, v in res.items() if int(v.get("STACK", 0)) > 0)
    print("\nkernels with a stack frame (register spills under the per-flavour register cap): %d of %d, largest %d bytes" %
          (len(spill), len(res), spill[-1][0] if spill else 0))
    print("HMMA / UTC*MMA = 0 everywhere by design: max-plus over integers has no tensor-core form (DESIGN.md section 4.1).")


if __name__ == "__main__":
    main()
