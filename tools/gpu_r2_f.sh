#!/bin/bash
# e2e sweep: contexts x pictures per step
TAG=${1:-r2f}
OUT=gpurun_out; mkdir -p $OUT
for v in "6 8" "8 8" "12 12" "8 16" "12 24" "16 16" "4 8"; do
  set -- $v
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-other --sustain 0 --no-verify --e2e-ctx $1 --e2e-pics $2 > $OUT/bench_${TAG}_c$1p$2.json 2> $OUT/bench_${TAG}_c$1p$2.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_c$1p$2.json")); e=d["e2e"]
    print("ctx $1 pics $2", "e2e", e["value"], "ms/pic", round(e["ms_per_step"]/e["pics_per_step_per_gpu"],4), "frac", e["pcie_frac"], e["achieved_gbs"], "drained", e["drained_step_value"], e["timing"][:110])
except Exception as ex:
    print("$v failed", ex); print(open("$OUT/bench_${TAG}_c$1p$2.err").read()[-400:])
PY
done
