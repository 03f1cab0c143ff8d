#!/bin/bash
# Builds libclv_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libclv_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v"
mkdir -p build
pids=()
for f in *.cu; do
  ( $NVCC $FLAGS -c "$f" -o "build/${f%.cu}.o" > "build/${f%.cu}.log" 2>&1 || { cat "build/${f%.cu}.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT build/*.o -lcudart
echo "built $(readlink -f $OUT)"
