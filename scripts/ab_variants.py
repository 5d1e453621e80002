#!/usr/bin/env python3
"""A/B harness for kernel variants (tuning aid): for every bialign_b200/build/variants/lib_<name>.so, copy it over the
in-tree library and run bench.py in a fresh process; prints value / e2e per variant.  Meant to run on the GPU box.
`--script <file> <args...>` runs that script instead of bench.py and prints its output lines as they are."""
import glob, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bialign_b200", "libbialign_b200.so")
script = None
argv = sys.argv[1:]
if len(argv) >= 2 and argv[0] == "--script":
    script, argv = argv[1], argv[2:]
extra = argv or ["--steps", "3", "--warmup", "3", "--no-cpu-baseline"]
for so in sorted(glob.glob(os.path.join(ROOT, "bialign_b200", "build", "variants", "lib_*.so"))):
    name = os.path.basename(so)[4:-3]
    shutil.copyfile(so, LIB)
    if script:
        out = subprocess.run([sys.executable, os.path.join(ROOT, script)] + extra, capture_output=True, text=True)
        for line in out.stdout.strip().splitlines():
            print(f"{name:12s} {line[:330]}", flush=True)
        if out.returncode:
            print(name, "FAILED", out.stderr[-400:], flush=True)
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"{name:12s} value {d['value']:.1f} e2e {d['e2e']['value']:.1f} ms/step {d['ms_per_step']:.2f} "
              f"frac {d['roofline']['frac']:.3f} sm_mhz {d['clocks']['sm_mhz']}", flush=True)
    except Exception as ex:
        print(name, "FAILED", ex, out.stderr[-400:], flush=True)
