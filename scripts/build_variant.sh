#!/bin/bash
# Build an A/B variant of libexb.so with extra -D flags:  scripts/build_variant.sh NAME -DEXB_ROW_STREAM=0 ...
# -> build/libexb_NAME.so   (select it at run time with EXB_LIB=build/libexb_NAME.so)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build/variant_$name
mkdir -p "$out"
pids=()
for src in "$root"/exponax_b200/csrc/*.cu; do
  obj=$out/$(basename "${src%.cu}").o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I"$root/include" "$@" -c "$src" -o "$obj" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -shared -o "$root/build/libexb_$name.so" "$out"/*.o
echo "built $root/build/libexb_$name.so"
