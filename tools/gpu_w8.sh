#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
export P265_LIB=$PWD/build_ab/lib_d.so
python tools/kbench.py --pics 16 --reps 30 --only config2 2>&1 | tee -a $OUT/kbench_c2.log
for rep in 1 2; do for v in r1 d b h t u v w; do
  export P265_LIB=$PWD/build_ab/lib_$v.so
  echo "== $v" | tee -a $OUT/kbench_c2.log
  P265_KB_MIX_ONLY=1 python tools/kbench.py --pics 16 --reps 30 --only config2 2>&1 | tee -a $OUT/kbench_c2.log
done; done
