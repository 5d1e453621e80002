#!/bin/bash
# scripts/build_variant.sh <name> [nvcc flags...]: build the library with extra nvcc flags into bialign_b200/build/variants/lib_<name>.so
# (the A/B harness scripts/ab_variants.py runs bench.py once per variant on the GPU box); the in-tree library is rebuilt without flags last.
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p bialign_b200/build/variants
BA_NVCC_EXTRA="$*" python -m bialign_b200.build --force > /dev/null
cp bialign_b200/libbialign_b200.so bialign_b200/build/variants/lib_$name.so
echo "built variant $name: $*"
