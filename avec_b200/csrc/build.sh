#!/bin/bash
# Build libavec_b200.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
# One object per source, compiled in parallel and only when the source (or a header) is newer than its object.
# Diagnostics build for tools/ts_probe.py: OUT=../libavec_b200_timeline.so OBJ=build_tl build.sh -DAVEC_TIMELINE
# (run with AVEC_LIB=<that file>)
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${OUT:-../libavec_b200.so}
OBJ=${OBJ:-build}
LOG=${LOG:-$OBJ/build.log}
SRCS="api.cu gemm_simt.cu gemm_tc.cu attention.cu attention_long.cu attention_mma.cu attention_tc.cu norm.cu convmod.cu frontend.cu ctc.cu train.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math -Xcompiler -fPIC -Xptxas -v $*"
mkdir -p $OBJ
# a change of flags rebuilds everything
if [ "$(cat $OBJ/.flags 2>/dev/null)" != "$FLAGS" ]; then rm -f $OBJ/*.o; echo "$FLAGS" > $OBJ/.flags; fi
pids=()
for s in $SRCS; do
  [ -f $s ] || continue
  o=$OBJ/${s%.cu}.o
  newest=$(ls -t $s *.cuh ../../include/avec_b200.h 2>/dev/null | head -1)
  if [ ! -f $o ] || [ $newest -nt $o ]; then
    ( $NVCC $FLAGS -c $s -o $o > $OBJ/${s%.cu}.log 2>&1 || { cat $OBJ/${s%.cu}.log; rm -f $o; exit 1; } ) &
    pids+=($!)
  fi
done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
[ $fail = 0 ] || { echo "build failed"; exit 1; }
cat $OBJ/*.log > $LOG 2>/dev/null || true
objs=""
for s in $SRCS; do [ -f $OBJ/${s%.cu}.o ] && objs="$objs $OBJ/${s%.cu}.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC $objs -o $OUT
grep -E "error|warning : .*spill|bytes spill" $LOG | grep -v "0 bytes spill" | head -20 || true
echo "built $OUT"
