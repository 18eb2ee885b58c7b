#!/usr/bin/env bash
# tools/build_variant_all.sh <name> [-Dflags...]: rebuild EVERY translation unit with extra flags into tune/libsphb_<name>.so
# (A/B tuning runs via SPHB_LIB, tools/tune_run.sh) — for macros that span several units
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="$ROOT/sph-particle-simulator_b200/csrc"
name=$1; shift 1
mkdir -p "$ROOT/tune/obj"
objs=(); pids=()
for f in api scan_sort neighbor pair pair_mask pair_mask_wide pair_stage pair_split integrate slab multi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 -I"$ROOT/include" -I"$SRC" -ccbin /usr/bin/g++ \
       "$@" -Xptxas -v -c "$SRC/$f.cu" -o "$ROOT/tune/obj/${f}_$name.o" 2> "$ROOT/tune/obj/${f}_$name.log" &
  pids+=($!)
  objs+=("$ROOT/tune/obj/${f}_$name.o")
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o "$ROOT/tune/libsphb_$name.so" "${objs[@]}"
echo "built tune/libsphb_$name.so"
