#!/bin/bash
# One GPU-box session: smoke, GPU parity tests, bench, then (only if the plain bench
# exited 0) the ncu launch list and one full capture of the two hot kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
BENCH_RC=$?
echo "bench exit $BENCH_RC"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
if [ "$BENCH_RC" = "0" ] && [ "$2" != "noncu" ]; then
  NCU_CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-other --e2e-steps 3 --e2e-pics 1"
  $NCU_CMD > $OUT/ncu_plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv $NCU_CMD > $OUT/ncu_launches_$TAG.log 2>&1
  echo "ncu launches exit $?"
  $NCU_CMD > $OUT/ncu_plain2_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'expand_kernel|residual_kernel|sao_kernel' -s 18 -c 6 -f -o $OUT/prof_$TAG $NCU_CMD > $OUT/ncu_full_$TAG.log 2>&1
  echo "ncu full exit $?"
fi
ls -la $OUT
