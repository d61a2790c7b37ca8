#!/bin/bash
# last GPU seconds of the round: launch list at HEAD, then (if time is left) a full capture of one step
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-other --no-verify --sustain 0 --e2e-steps 3 --e2e-pics 1"
timeout 50 $CMD > $OUT/plain_fin.log 2>&1 &&
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_fin.csv $CMD > $OUT/ncu_launches_fin.log 2>&1
echo "launch list exit $?"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:'expand_kernel|residual_kernel|sao_kernel' -s 18 -c 6 -f -o $OUT/prof_fin $CMD > $OUT/ncu_full_fin.log 2>&1
echo "full capture exit $?"
