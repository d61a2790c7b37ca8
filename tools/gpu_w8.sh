#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2 3; do for v in r1 d v u; do
  export P265_LIB=$PWD/build_ab/lib_$v.so
  echo "== $v" | tee -a $OUT/kbench_c2b.log
  P265_KB_MIX_ONLY=1 python tools/kbench.py --pics 8 --reps 40 --only config2 2>&1 | tee -a $OUT/kbench_c2b.log
  P265_KB_MIX_ONLY=1 python tools/kbench.py --pics 16 --reps 30 --only residual --quick 2>&1 | tee -a $OUT/kbench_c2b.log
done; done
