#!/bin/bash
# ncu capture of the big bins with and without zero-extent codes (tools/kbench.py --only zprof)
TAG=${1:-zprof}
OUT=gpurun_out; mkdir -p $OUT
KB="python tools/kbench.py --only zprof --pics 16 --reps 2"
$KB > $OUT/kb_$TAG.log 2>&1 && cat $OUT/kb_$TAG.log &&
ncu --set full --clock-control none --import-source on -k regex:'residual_kernel' -c 20 -f -o $OUT/prof_$TAG $KB > $OUT/ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_$TAG.log
