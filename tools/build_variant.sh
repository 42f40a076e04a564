#!/bin/bash
# Builds a variant of libapx.so with extra -D flags for ONE source file into variants/NAME/ (A/B runs on the GPU box:
# LD_LIBRARY_PATH=variants/NAME tools/apx_harness ... ; the harness finds libapx through a RUNPATH, which LD_LIBRARY_PATH overrides).
#   tools/build_variant.sh NAME file.cu "-DFOO=1 -DBAR=2"
set -e
cd "$(dirname "$0")/../ap_adapter_b200/csrc"
NAME=$1; SRC=$2; FLAGS=$3
OUT=../../variants/$NAME; mkdir -p $OUT
make -j8 > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $FLAGS -c $SRC -o $OUT/${SRC%.cu}.o 2> $OUT/ptxas.log
grep -E "spill" $OUT/ptxas.log | sort | uniq -c
OBJS=$(ls build/*.o | grep -v "build/${SRC%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libapx.so $OBJS $OUT/${SRC%.cu}.o
ls -la $OUT/libapx.so
