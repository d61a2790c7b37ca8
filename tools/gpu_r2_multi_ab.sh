#!/bin/bash
# N-GPU A/B of the e2e transport on one box: default / no zero-copy / round-1 dense / fewer contexts
N=${1:-8}; TAG=${2:-r2n}
OUT=gpurun_out; mkdir -p $OUT
run() { # name env args...
  NAME=$1; ENVV=$2; shift 2
  env $ENVV timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 6 --warmup 3 --no-other --no-verify --sustain 0 "$@" > $OUT/bench_${TAG}_${N}_$NAME.json 2> $OUT/bench_${TAG}_${N}_$NAME.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_${N}_$NAME.json")); e=d["e2e"]
    print("N=$N %-8s" % "$NAME", "e2e", e["value"], "ms", e["ms_per_step"], "probe sum", e["pcie"]["sum_over_ranks_gbs"], "frac", e["pcie_frac"], "achieved/rank", e["achieved_gbs"], "drained", e["drained_step_value"])
except Exception as ex:
    print("$NAME failed", ex); print(open("$OUT/bench_${TAG}_${N}_$NAME.err").read()[-500:])
PY
}
run default "A=1"
run nozc "P265_NO_ZERO_COPY=1"
run dense "A=1" --e2e-dense
run ctx3 "A=1" --e2e-ctx 3
run ctx2pics4 "A=1" --e2e-ctx 2 --e2e-pics 4
