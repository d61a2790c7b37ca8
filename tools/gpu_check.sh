#!/bin/bash
# smoke + the whole GPU suite + one bench line:  bash tools/gpu_check.sh tag
TAG=${1:-chk}
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1700 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc $?"
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["kernels"], "e2e", d["e2e"]["value"])
for k,v in d["other_kernels"].items(): print(k, {a:b for a,b in v.items() if a!="what"})
print(d.get("verify")); print(d.get("sustained"))
PY
