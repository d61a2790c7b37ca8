#!/bin/bash
# A/B a compile-time variant on the GPU box:  bash tools/gpu_ab.sh tag "<nvcc extra>" [bench args]
TAG=$1; EXTRA=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
P265_NVCC_EXTRA="$EXTRA" python -m p265_b200.build --force > $OUT/build_$TAG.log 2>&1 || { echo "build failed"; tail -5 $OUT/build_$TAG.log; exit 1; }
python bench.py --no-cpu "$@" > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench $TAG exit $?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$TAG.json"))
    print("$TAG", "value", d["value"], "ms/step", d["ms_per_step"], {k:v["ms"] for k,v in d["roofline"]["kernels"].items()}, "e2e", d["e2e"]["value"], d["clocks"])
except Exception as e:
    print("$TAG parse error", e); print(open("$OUT/bench_$TAG.err").read()[-800:])
PY
