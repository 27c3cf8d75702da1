#!/bin/bash
# Build libavec_b200.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
# Diagnostics build for tools/ts_probe.py: OUT=../libavec_b200_timeline.so build.sh -DAVEC_TIMELINE (run with AVEC_LIB=<that file>)
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${OUT:-../libavec_b200.so}
SRCS="api.cu gemm_simt.cu gemm_tc.cu attention.cu attention_long.cu attention_mma.cu norm.cu convmod.cu frontend.cu ctc.cu train.cu"
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math -Xcompiler -fPIC -shared \
  -Xptxas -v "$@" $SRCS -o $OUT 2> ${LOG:-build.log} || { cat ${LOG:-build.log}; exit 1; }
grep -E "error|warning : .*spill|bytes spill" ${LOG:-build.log} | grep -v "0 bytes spill" | head -20 || true
echo "built $OUT"
