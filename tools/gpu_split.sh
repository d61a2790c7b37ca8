#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
python tools/kbench.py --only residual --quick --pics 16 --reps 20 2>&1 | head -1 | tee -a $OUT/kbench_split.log
for p in $1; do echo "== P265_GRID_PCT=$p" | tee -a $OUT/kbench_split.log; P265_GRID_PCT=$p python tools/kbench.py --only split --pics 16 --reps 20 2>&1 | tee -a $OUT/kbench_split.log; done
