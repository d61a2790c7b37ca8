#!/bin/bash
# Zero-aware passes: parity (new tests + the whole GPU suite), then A/B against a build without the
# dispatch (build_ab/lib_nozext.so, tools/build_variant.sh nozext "-DP265_ZERO_EXTENT=0") and the lowfreq timings.
TAG=${1:-r2z}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_extents.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu_$TAG.log
for rep in 1 2; do for v in nozext new; do
  if [ $v = new ]; then unset P265_LIB; else export P265_LIB=$PWD/build_ab/lib_$v.so; fi
  echo "== $v" | tee -a $OUT/kbench_$TAG.log
  python tools/kbench.py --pics 16 --reps 30 --only residual --quick 2>&1 | tee -a $OUT/kbench_$TAG.log
done; done
unset P265_LIB
python tools/kbench.py --pics 16 --reps 30 --only lowfreq 2>&1 | tee -a $OUT/kbench_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value", d["value"], d["roofline"]["kernels"], "e2e", d["e2e"]["value"])
print(d["other_kernels"].get("zero_aware_residual_4k10_lowfreq"))
print(d.get("verify"))
PY
