#!/bin/bash
# Round-end evidence on the GPU box (run through gpurun): tests, the bench line of both arms, the per-config table, then
# -- only after the plain runs exited 0 -- the ncu launch list and one full capture of the hot kernel.
# usage: [SKIP_NCU=1 | ONLY_NCU=1] scripts/capture_round.sh <tag>   (writes gpurun_out/*_<tag>.*)
set -u
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
if [ -z "${ONLY_NCU:-}" ]; then
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $out/pytest_$tag.log
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err || { echo "bench failed"; tail -5 $out/bench_$tag.err; exit 1; }
tail -c 600 $out/bench_$tag.json
timeout 600 python bench.py --impl reference > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err || echo "reference arm failed"
timeout 600 python scripts/bench_configs.py > $out/per_config_$tag.jsonl 2>> $out/bench_$tag.err || echo "bench_configs failed"
fi
[ -n "${SKIP_NCU:-}" ] && exit 0
small="bench.py --steps 1 --warmup 1 --no-cpu-baseline --pairs-per-gpu 592"
timeout 300 python $small > $out/plain_$tag.log 2>&1 || { echo "small bench failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv python $small > $out/ncu1_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fill_systolic -c 1 -o $out/prof_$tag python $small > $out/ncu2_$tag.log 2>&1
ls -la $out | tail -12
