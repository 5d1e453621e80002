#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (s-will/BiAlign, Cython) into
# oracle/_ref/ (git-ignored, travels to the GPU box).  Nothing in the product path imports it.
# Sources are compiled where they lie under /root/reference; only build OUTPUTS land in oracle/_ref
# (the two extension modules; the generated C is removed again).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -f "$REF/src/bialignment.pyx" ] || { echo "reference not present at $REF (expected on the GPU box); keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT"
PY=${PYTHON:-python3}
EXT=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
INC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
if [ "$OUT/bialignment$EXT" -nt "$REF/src/bialignment.pyx" ] && [ "$OUT/bialignment_nonpyx$EXT" -nt "$REF/src/bialignment_nonpyx.py" ]; then
    echo "oracle/_ref up to date"; exit 0
fi
# same directives as the reference's setup.py:13-18
$PY -m cython -3 -X boundscheck=False "$REF/src/bialignment.pyx" -o "$OUT/bialignment.c"
gcc -O2 -fPIC -shared -fwrapv -fno-strict-aliasing -I"$INC" "$OUT/bialignment.c" -o "$OUT/bialignment$EXT"
# the pure-python helper module the extension imports at run time is compiled too (Cython accepts .py), so that
# oracle/_ref holds binaries only -- no copy of a reference source file
# (--lenient: the module names `sys` on an error path without importing it; as in the interpreted module that stays
# a run-time NameError on that path only)
$PY -m cython -3 --lenient "$REF/src/bialignment_nonpyx.py" -o "$OUT/bialignment_nonpyx.c"
gcc -O2 -fPIC -shared -fwrapv -fno-strict-aliasing -I"$INC" "$OUT/bialignment_nonpyx.c" -o "$OUT/bialignment_nonpyx$EXT"
rm -f "$OUT/bialignment.c" "$OUT/bialignment_nonpyx.c" "$OUT/bialignment_nonpyx.py"
echo "built $OUT/bialignment$EXT"
