#!/bin/bash
# A/B of prebuilt libraries on the GPU box (tools/build_variant.sh):
#   bash tools/gpu_ab2.sh tag "base new b0c10" [kbench args]     ("new" = the in-tree library)
TAG=$1; LIBS=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
if [ -n "$AB_TESTS" ]; then
timeout 600 python -m pytest tests/test_gpu_residual.py tests/test_gpu_dropin.py tests/test_gpu_decode_sanity.py tests/test_gpu_fuzz_streams.py -q -x -p no:cacheprovider > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_$TAG.log
fi
for rep in 1 2; do for v in $LIBS; do
  if [ $v = new ]; then unset P265_LIB; else export P265_LIB=$PWD/build_ab/lib_$v.so; fi
  echo "== $v" | tee -a $OUT/kbench_$TAG.log
  python tools/kbench.py "$@" 2>&1 | tee -a $OUT/kbench_$TAG.log
done; done
