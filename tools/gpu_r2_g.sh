#!/bin/bash
# e2e: residual / SAO calls on separate contexts, CUDA_DEVICE_MAX_CONNECTIONS
TAG=${1:-r2g}
OUT=gpurun_out; mkdir -p $OUT
run() { NAME=$1; ENVV=$2; shift 2
  env $ENVV timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-other --sustain 0 --no-verify "$@" > $OUT/bench_${TAG}_$NAME.json 2> $OUT/bench_${TAG}_$NAME.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_$NAME.json")); e=d["e2e"]
    print("%-22s" % "$NAME", "e2e", e["value"], "ms/pic", round(e["ms_per_step"]/e["pics_per_step_per_gpu"],4), "frac", e["pcie_frac"], e["achieved_gbs"], "drained", e["drained_step_value"])
except Exception as ex:
    print("$NAME failed", ex); print(open("$OUT/bench_${TAG}_$NAME.err").read()[-400:])
PY
}
run same6 A=1 --no-e2e-split
run split6 A=1
run split8 A=1 --e2e-ctx 8
run split12 A=1 --e2e-ctx 12 --e2e-pics 12
run split8_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 --e2e-ctx 8
run same6_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 --no-e2e-split
run split16_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32 --e2e-ctx 16 --e2e-pics 16
run split4 A=1 --e2e-ctx 4
