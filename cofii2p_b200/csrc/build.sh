#!/bin/bash
# Builds libcofi_b200.so for sm_100a in-tree (the .so travels to the GPU box with the gpurun snapshot).
set -e
cd "$(dirname "$0")"
OUT=../libcofi_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p ../../build/obj
objs=""
for f in *.cu; do
  o=../../build/obj/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ ../../include/cofi_b200.h -nt "$o" ]; then
    echo "nvcc $f"
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
  fi
  objs="$objs $o"
done
wait
$NVCC -shared -o $OUT $objs -lcudart
echo "built $(realpath $OUT)"
