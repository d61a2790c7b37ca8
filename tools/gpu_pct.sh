#!/bin/bash
# occupancy sensitivity of the persistent residual grids: bash tools/gpu_pct.sh "100 75 50"
OUT=gpurun_out; mkdir -p $OUT
for p in $1; do echo "== P265_GRID_PCT=$p" | tee -a $OUT/kbench_pct.log; P265_GRID_PCT=$p python tools/kbench.py --only residual --quick --pics 16 --reps 20 2>&1 | tee -a $OUT/kbench_pct.log; done
