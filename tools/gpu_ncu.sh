#!/bin/bash
# ncu full capture of the hot kernels:  bash tools/gpu_ncu.sh tag "<kernel regex>" [nvcc extra]
TAG=$1; RE=$2; EXTRA=$3
OUT=gpurun_out; mkdir -p $OUT
if [ -n "$EXTRA" ]; then P265_NVCC_EXTRA="$EXTRA" python -m p265_b200.build --force > $OUT/build_$TAG.log 2>&1 || exit 1; fi
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 3 --e2e-pics 1"
$CMD > $OUT/ncu_plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 6 -c 2 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_full_$TAG.log | cut -c1-300
