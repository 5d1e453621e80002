#!/bin/bash
# One full-wave ncu capture of the hot fill kernel (run through gpurun):  scripts/capture_ncu.sh <tag> [extra bench args]
# 3552 pairs = 8 work items per resident CTA (444 CTAs), so the capture is not dominated by the tail of the wave.
set -u
tag=${1:-rX}; shift || true
out=gpurun_out
mkdir -p $out
small="bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-shapes --pairs-per-gpu 3552 $*"
timeout 300 python $small > $out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 $out/plain_$tag.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv python $small > $out/ncu1_$tag.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fill_systolic -c 1 -o $out/prof_$tag python $small > $out/ncu2_$tag.log 2>&1
ls -la $out/prof_$tag.ncu-rep
