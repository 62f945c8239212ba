#!/bin/bash
# Build a variant of the library for kernel A/B timing:  tools/build_variant.sh NAME "-DFOO=1 ..."  ->  variants/libcobevt_NAME.so
set -e
cd "$(dirname "$0")/../cobevt_b200/csrc"
OUT=../../variants; mkdir -p $OUT
TMP=$(mktemp -d)
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $2"
for f in *.cu; do nvcc $FLAGS -c "$f" -o "$TMP/${f%.cu}.o" & done; wait
nvcc -shared -o $OUT/libcobevt_$1.so $TMP/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf $TMP; echo "built variants/libcobevt_$1.so"
