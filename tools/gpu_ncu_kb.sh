#!/bin/bash
# full ncu capture of the residual bin kernels from tools/kbench.py: bash tools/gpu_ncu_kb.sh tag [lib]
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
[ -n "$2" ] && export P265_LIB=$PWD/build_ab/lib_$2.so
KB="python tools/kbench.py --only residual --quick --pics 16 --reps 2"
$KB > $OUT/kb_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/kb_plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'residual_kernel' -s 20 -c 10 -f -o $OUT/prof_kb_$TAG $KB > $OUT/ncu_kb_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_kb_$TAG.log
