#!/bin/bash
# Persistent grid size per bin (P265_GRID_PCT_BINS, per cent of the resident capacity; bins 32,16,8,4) under the
# least-work-first launch order
TAG=${1:-wave2}
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { echo "== $1" | tee -a $OUT/kbench_$TAG.log; env $1 P265_KB_MIX_ONLY=1 python tools/kbench.py --pics 16 --reps 30 --only residual --quick 2>&1 | tee -a $OUT/kbench_$TAG.log; }
run A=1
run P265_GRID_PCT_BINS=100,100,100,50
run P265_GRID_PCT_BINS=100,100,100,75
run P265_GRID_PCT_BINS=100,100,100,90
run P265_GRID_PCT_BINS=90,100,100,100
run P265_GRID_PCT_BINS=100,90,100,100
run P265_GRID_PCT_BINS=100,100,90,100
run P265_GRID_PCT_BINS=95,95,95,95
run A=2
