#!/bin/bash
# Diagnostic: small-bin TBs in raster order inside a plane (P265_KB_RASTER=1) against decoding order
TAG=${1:-raster}
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do for r in 1 2; do
  echo "== raster=$r" | tee -a $OUT/kbench_${TAG}b.log
  P265_KB_RASTER=$r python tools/kbench.py --pics 16 --reps 30 --only residual --quick 2>&1 | tee -a $OUT/kbench_${TAG}b.log
  P265_KB_RASTER=$r python tools/kbench.py --pics 8 --reps 30 --only config2 2>&1 | tee -a $OUT/kbench_${TAG}b.log
done; done
