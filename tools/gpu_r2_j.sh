#!/bin/bash
TAG=${1:-r2j}
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1700 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider 2>&1 | tail -4
python tools/kbench.py --pics 16 --reps 30 --only residual --quick 2>&1 | tee $OUT/kbench_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-verify > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value", d["value"], d["roofline"]["kernels"], "c2", d["other_kernels"]["config2_residual_1080p8"], "e2e", d["e2e"]["value"])
PY
