#!/bin/bash
# bin launch order in the contract benchmark's step (residual chain + SAO)
OUT=gpurun_out; mkdir -p $OUT
for o in 0123 3012 3102 0123 3012 1023; do
  P265_BIN_ORDER=$o timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-other --no-verify --sustain 0 --e2e-steps 2 > $OUT/bench_order_$o.json 2> $OUT/bench_order_$o.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_order_$o.json"))
print("$o", "value", d["value"], "ms/step", d["ms_per_step"], {k:v["ms"] for k,v in d["roofline"]["kernels"].items()})
PY
done
