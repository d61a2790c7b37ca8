#!/bin/bash
# two-chain residual launch (P265_SPLIT) and residual || SAO overlap (P265_GRID_PCT)
TAG=${1:-r2e}
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/split_$TAG.log; : > $L
export P265_KB_MIX_ONLY=1
run() { echo "== $*" | tee -a $L; env "$@" python tools/kbench.py --pics 16 --reps 30 --only residual --quick 2>&1 | tee -a $L; }
run A=1
run P265_SPLIT="03|12" P265_SPLIT_PCT=70,70
run P265_SPLIT="03|12" P265_SPLIT_PCT=75,50
run P265_SPLIT="03|12" P265_SPLIT_PCT=50,75
run P265_SPLIT="03|12" P265_SPLIT_PCT=100,100
run P265_SPLIT="01|23" P265_SPLIT_PCT=75,50
run P265_SPLIT="01|23" P265_SPLIT_PCT=75,100
run P265_SPLIT="01|23" P265_SPLIT_PCT=100,100
run P265_SPLIT="02|13" P265_SPLIT_PCT=75,75
run P265_SPLIT="0|123" P265_SPLIT_PCT=75,75
unset P265_KB_MIX_ONLY
for pct in 100 75 50; do echo "== overlap GRID_PCT $pct" | tee -a $L; P265_GRID_PCT=$pct python tools/kbench.py --pics 16 --reps 30 --only overlap 2>&1 | tee -a $L; done
P265_SPLIT="03|12" P265_SPLIT_PCT=70,70 timeout 600 python -m pytest tests/test_gpu_residual.py tests/test_gpu_transport.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
