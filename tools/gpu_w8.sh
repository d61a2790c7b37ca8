#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do for v in base d j m n o p q r s; do
  export P265_LIB=$PWD/build_ab/lib_$v.so
  echo "== $v" | tee -a $OUT/kbench_w8d.log
  P265_KB_MIX_ONLY=1 python tools/kbench.py --pics 16 --reps 40 --only residual --quick 2>&1 | tee -a $OUT/kbench_w8d.log
done; done
