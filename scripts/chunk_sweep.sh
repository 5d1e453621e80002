#!/bin/bash
# cfg5 (one 8192 x 8192 pair) with the row blocks cut into BA_COL_CHUNKS column chunks (tuning aid; run through gpurun)
for c in 1 8 14 24; do
  export BA_COL_CHUNKS=$c
  echo "COL_CHUNKS=$c"; timeout 120 python scripts/long_ab.py 5 2>&1 | grep "io_warp=1" | grep -o '"config": "[^"]*"\|"fill_ms": [0-9.]*\|"gcups_fill": [0-9.]*\|"total_ms": [0-9.]*' | paste - - - -
done
