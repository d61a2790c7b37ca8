#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do for v in d i2 f2 m2 n2 o2 p2 q2; do
  export P265_LIB=$PWD/build_ab/lib_$v.so
  echo "== $v" | tee -a $OUT/kbench_w8f.log
  P265_KB_MIX_ONLY=1 python tools/kbench.py --pics 16 --reps 30 --only residual --quick 2>&1 | tee -a $OUT/kbench_w8f.log
  P265_KB_MIX_ONLY=1 python tools/kbench.py --pics 8 --reps 30 --only config2 2>&1 | tee -a $OUT/kbench_w8f.log
done; done
