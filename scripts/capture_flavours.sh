#!/bin/bash
# ncu --set full captures of the kernels other than the headline one (run through gpurun):
#   long-pair flavour (config 5), 16-bit pair mode (config 4), dedicated non-affine kernel, traceback kernel.
# usage: scripts/capture_flavours.sh <tag>      (writes gpurun_out/prof_<tag>_<flavour>.ncu-rep)
set -u
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 900 $NCU -k regex:fill_systolic -o $out/prof_${tag}_long python scripts/bench_configs.py 5 > $out/ncu_${tag}_long.log 2>&1
timeout 900 $NCU -k regex:fill_systolic -o $out/prof_${tag}_p16 python scripts/bench_configs.py 4s > $out/ncu_${tag}_p16.log 2>&1
timeout 900 $NCU -k regex:fill_na -o $out/prof_${tag}_na python scripts/bench_configs.py na > $out/ncu_${tag}_na.log 2>&1
timeout 900 $NCU -k regex:traceback -o $out/prof_${tag}_traceback python scripts/bench_configs.py 3s > $out/ncu_${tag}_tb.log 2>&1
ls -la $out/prof_${tag}_*.ncu-rep
