#!/bin/bash
# Build libcobevt_sm100.so in-tree (sm_100a only; nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libcobevt_sm100.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
OBJS=""
for f in *.cu; do
  o="${f%.cu}.o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ sm100.cuh -nt "$o" ] || [ common.h -nt "$o" ] || [ ../../include/cobevt_b200.h -nt "$o" ]; then
    nvcc $FLAGS -c "$f" -o "$o" 2> "${f%.cu}.ptxas.log" || { cat "${f%.cu}.ptxas.log"; exit 1; }
  fi
  OBJS="$OBJS $o"
done
nvcc -shared -o $OUT $OBJS -gencode arch=compute_100a,code=sm_100a -lcudart
echo "built $OUT"
