#!/bin/bash
# Final evidence run: launch list + full ncu capture of one step (4 residual bin kernels +
# SAO), after the same command has exited 0 without ncu.   bash tools/gpu_profile_final.sh tag
TAG=${1:-r1}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 3 --e2e-pics 1"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'residual_kernel|sao_kernel' -s 15 -c 5 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
