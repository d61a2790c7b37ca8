#!/bin/bash
# Multi-GPU session: N = number of GPUs of this box.  Pool tests, torchrun bench (as the driver launches it),
# single-process pool e2e.
N=${1:-2}; TAG=${2:-r2m}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_${TAG}_$N.txt 2>&1
timeout 600 python -m pytest tests/test_pool.py tests/test_multi_rank.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-other > $OUT/bench_${TAG}_$N.json 2> $OUT/bench_${TAG}_$N.err; echo "bench N=$N exit $?"
tail -c 300 $OUT/bench_${TAG}_$N.err
python - <<PY
import json
d=json.load(open("$OUT/bench_${TAG}_$N.json")); e=d["e2e"]
print("N=$N value", d["value"], "e2e", e["value"], "ms", e["ms_per_step"], "pcie", e["pcie"], "frac", e["pcie_frac"], e["achieved_gbs"], "drained", e["drained_step_value"])
PY
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-other --no-cpu --sustain 0 --e2e-pool --e2e-pics $((8*N)) > $OUT/bench_${TAG}_pool$N.json 2> $OUT/bench_${TAG}_pool$N.err; echo "pool exit $?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_pool$N.json")); e=d["e2e"]
    print("pool over $N GPUs in one process: e2e", e["value"], "ms", e["ms_per_step"], e["pool"], e["timing"])
except Exception as ex:
    print("pool failed", ex); print(open("$OUT/bench_${TAG}_pool$N.err").read()[-800:])
PY
