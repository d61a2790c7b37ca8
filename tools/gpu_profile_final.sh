#!/bin/bash
# Final evidence run: launch list + full ncu capture of one step (4 residual bin kernels +
# SAO), after the same command has exited 0 without ncu; then the neighbouring kernels
# (deblocking, reconstruction) from tools/kbench.py.   bash tools/gpu_profile_final.sh tag
TAG=${1:-r1}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-other --e2e-steps 3 --e2e-pics 1"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'expand_kernel|residual_kernel|sao_kernel' -s 18 -c 6 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
KB="python tools/kbench.py --only deblock,recon --pics 8 --reps 3"
$KB > $OUT/kb_plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'deblock_kernel|recon_kernel' -c 12 -f -o $OUT/prof_other_$TAG $KB > $OUT/ncu_other_$TAG.log 2>&1
echo "other kernels capture exit $?"
python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err; echo "ref rc $?"
