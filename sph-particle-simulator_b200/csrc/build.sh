#!/usr/bin/env bash
# Builds libsphb.so (CUDA kernels + C ABI) for sm_100a, in-tree.  No GPU needed to build.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/../libsphb.so"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 -I"$ROOT/include" -I"$HERE"
       -ccbin /usr/bin/g++ ${SPHB_NVCC_EXTRA:-})
OBJ="$HERE/build"
mkdir -p "$OBJ"
pids=()
for f in api scan_sort neighbor pair pair_mask pair_mask_wide pair_stage pair_split integrate slab multi; do
  if [ ! -f "$OBJ/$f.o" ] || [ "$HERE/$f.cu" -nt "$OBJ/$f.o" ] || [ -n "$(find "$HERE" "$ROOT/include" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$OBJ/$f.o" 2>/dev/null)" ]; then
    "$NVCC" "${FLAGS[@]}" -Xptxas -v -c "$HERE/$f.cu" -o "$OBJ/$f.o" 2> "$OBJ/$f.ptxas.log" &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do [ -z "$p" ] || wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "$OBJ"/*.ptxas.log | grep -E 'error|Error' -A3 >&2 || true; exit 1; fi
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o "$OUT" "$OBJ"/api.o "$OBJ"/scan_sort.o "$OBJ"/neighbor.o "$OBJ"/pair.o "$OBJ"/pair_mask.o "$OBJ"/pair_mask_wide.o "$OBJ"/pair_stage.o "$OBJ"/pair_split.o "$OBJ"/integrate.o "$OBJ"/slab.o "$OBJ"/multi.o
echo "built $OUT"
