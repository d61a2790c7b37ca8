#!/bin/bash
# e2e anatomy + small-bin streaming A/B
TAG=${1:-r2c}
OUT=gpurun_out; mkdir -p $OUT
python tools/e2e_probe2.py > $OUT/e2e_probe_$TAG.txt 2>&1; cat $OUT/e2e_probe_$TAG.txt
P265_NO_ZERO_COPY=1 N_PIC=24 python tools/e2e_probe2.py 2>&1 | grep -E "sao in place  |pipelined.*inplace=True" | sed 's/^/nozc: /' | tee -a $OUT/e2e_probe_$TAG.txt
for lib in stream persist stream10; do
  echo "== $lib" | tee -a $OUT/kbench_$TAG.log
  P265_LIB=$PWD/build_ab/lib_$lib.so python tools/kbench.py --pics 16 --reps 20 --only residual --quick 2>&1 | tee -a $OUT/kbench_$TAG.log
done
P265_LIB=$PWD/build_ab/lib_stream.so python tools/kbench.py --pics 32 --reps 10 --only residual --quick 2>&1 | sed 's/^/pics32: /' | tee -a $OUT/kbench_$TAG.log
P265_LIB=$PWD/build_ab/lib_stream.so timeout 600 python -m pytest tests/test_gpu_residual.py tests/test_gpu_transport.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
