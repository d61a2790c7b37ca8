#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for o in $1; do echo "== P265_BIN_ORDER=$o" | tee -a $OUT/kbench_order.log; P265_BIN_ORDER=$o python tools/kbench.py --only residual --quick --pics 16 --reps 20 2>&1 | grep mix | tee -a $OUT/kbench_order.log; done
