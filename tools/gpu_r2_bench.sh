#!/bin/bash
# Round-2 GPU session: GPU tests, then bench (packed transport), A/B legs of the e2e path.
TAG=${1:-r2b}
OUT=gpurun_out; mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -15 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
for v in dense nozc ctx3 ctx8 pics16; do
  case $v in
    dense) ARGS="--e2e-dense"; ENVV="" ;;
    nozc) ARGS=""; ENVV="P265_NO_ZERO_COPY=1" ;;
    ctx3) ARGS="--e2e-ctx 3"; ENVV="" ;;
    ctx8) ARGS="--e2e-ctx 8"; ENVV="" ;;
    pics16) ARGS="--e2e-pics 16 --e2e-ctx 8"; ENVV="" ;;
  esac
  env $ENVV timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-other --sustain 0 --no-verify $ARGS > $OUT/bench_${TAG}_$v.json 2> $OUT/bench_${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_$v.json")); e=d["e2e"]
    print("$v", "e2e", e["value"], "ms", e["ms_per_step"], "h2d", e["h2d_bytes_per_step"]/e["pics_per_step_per_gpu"]/1e6, "d2h", e["d2h_bytes_per_step"]/e["pics_per_step_per_gpu"]/1e6, "frac", e["pcie_frac"], e["achieved_gbs"], e["pcie"]["h2d_gbs"], e["pcie"]["d2h_gbs"], e["timing"])
except Exception as ex:
    print("$v failed", ex); print(open("$OUT/bench_${TAG}_$v.err").read()[-600:])
PY
done
