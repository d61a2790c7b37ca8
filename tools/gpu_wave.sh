#!/bin/bash
# Second-wave experiment: persistent grids larger than the resident capacity (P265_GRID_PCT > 100)
TAG=${1:-wave}
OUT=gpurun_out; mkdir -p $OUT
run() { echo "== $1" | tee -a $OUT/kbench_$TAG.log; env $1 P265_KB_MIX_ONLY=1 python tools/kbench.py --pics 16 --reps 40 --only residual --quick 2>&1 | tee -a $OUT/kbench_$TAG.log; }
run A=1
run P265_GRID_PCT=110
run P265_GRID_PCT=125
run P265_GRID_PCT=150
run P265_GRID_PCT=200
run P265_GRID_PCT_BINS=125,125,100,100
run P265_GRID_PCT_BINS=100,100,125,125
run P265_GRID_PCT_BINS=100,100,150,150
run P265_GRID_PCT_BINS=100,100,200,200
run P265_GRID_PCT_BINS=125,125,150,100
run A=2
