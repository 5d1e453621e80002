#!/bin/bash
# flag period of the long-pair pipeline (BA_LONG_LQ, in ring periods) on the two single pairs (tuning aid; run through gpurun)
for lq in ${LQS:-0 1 2 3 6}; do
  if [ $lq = 0 ]; then unset BA_LONG_LQ; else export BA_LONG_LQ=$lq; fi
  echo "LQ=$lq"; timeout 200 python scripts/long_ab.py 2 5 2>&1 | grep "io_warp=1" | grep -o '"config": "[^"]*"\|"fill_ms": [0-9.]*\|"gcups_fill": [0-9.]*' | paste - - -
done
