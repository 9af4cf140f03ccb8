#!/bin/bash
# Builds libcofi_b200.so for sm_100a in-tree (the .so travels to the GPU box with the gpurun snapshot).
# An object is rebuilt when its .cu, ANY header (*.cuh, include/*.h), this script or the flags changed: the state is
# a content hash per object (build/obj/<name>.hash), not a timestamp.  The stale object is deleted before compiling and
# every compile job's exit status is checked, so a failed compile can never link an old object into the library.
set -u
cd "$(dirname "$0")"
OUT=../libcofi_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Wno-deprecated-gpu-targets"
OBJ=../../build/obj
mkdir -p "$OBJ"
hdr_hash=$(cat *.cuh ../../include/*.h build.sh | sha256sum | cut -d' ' -f1)
objs=""
pids=()
names=()
for f in *.cu; do
  o=$OBJ/${f%.cu}.o
  h=$OBJ/${f%.cu}.hash
  want="$(sha256sum < "$f" | cut -d' ' -f1) $hdr_hash $FLAGS ${PTXAS_V:-}"
  if [ ! -f "$o" ] || [ ! -f "$h" ] || [ "$(cat "$h")" != "$want" ]; then
    echo "nvcc $f"
    rm -f "$o" "$h"
    ( $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" && echo "$want" > "$h" ) &
    pids+=($!)
    names+=("$f")
  fi
  objs="$objs $o"
done
fail=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then
    echo "build.sh: compiling ${names[$i]} FAILED" >&2
    fail=1
  fi
done
if [ $fail -ne 0 ]; then
  rm -f "$OUT"
  exit 1
fi
rm -f "$OUT"
$NVCC -Wno-deprecated-gpu-targets -shared -o $OUT $objs -lcudart || exit 1
echo "built $(realpath $OUT)"
