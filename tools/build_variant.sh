#!/bin/bash
# Build a tuning variant of the library into build_ab/lib_<name>.so:  bash tools/build_variant.sh name "<nvcc extra>"
# (only residual.cu is recompiled per variant; the other objects are shared)
NAME=$1; EXTRA=$2
mkdir -p build_ab/$NAME
for f in api residual coeffs transport sao recon deblock peak; do
  if [ $f = residual ] || [ ! -f build_ab/common_$f.o ] || [ p265_b200/csrc/$f.cu -nt build_ab/common_$f.o ]; then
    o=build_ab/$NAME/$f.o; [ $f != residual ] && o=build_ab/common_$f.o
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas=-v $EXTRA \
      -I include -I p265_b200/csrc -c p265_b200/csrc/$f.cu -o $o 2> build_ab/$NAME/$f.ptxas.log || { echo "nvcc failed for $f"; tail -5 build_ab/$NAME/$f.ptxas.log; exit 1; }
  fi
done
nvcc -shared -o build_ab/lib_$NAME.so build_ab/$NAME/residual.o build_ab/common_api.o build_ab/common_coeffs.o build_ab/common_transport.o build_ab/common_sao.o build_ab/common_recon.o build_ab/common_deblock.o build_ab/common_peak.o -gencode arch=compute_100a,code=sm_100a
grep -A2 "residual_kernelILi[0-3]ELi2" build_ab/$NAME/residual.ptxas.log | grep -E "Used|spill" | paste - - | sed 's/ptxas info    ://g' 
