#!/usr/bin/env bash
# tools/build_variant.sh <name> <file.cu> [-Dflags...]: rebuild ONE translation unit with extra flags and link it with the
# other (already built) objects into tune/libsphb_<name>.so — for A/B tuning runs via SPHB_LIB (tools/tune_run.sh)
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="$ROOT/sph-particle-simulator_b200/csrc"
name=$1; unit=$2; shift 2
mkdir -p "$ROOT/tune/obj"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 -I"$ROOT/include" -I"$SRC" -ccbin /usr/bin/g++ \
     "$@" -Xptxas -v -c "$SRC/$unit.cu" -o "$ROOT/tune/obj/${unit}_$name.o" 2> "$ROOT/tune/obj/${unit}_$name.log"
objs=()
for f in api scan_sort neighbor pair pair_mask pair_mask_wide pair_stage pair_split integrate slab multi; do
  if [ "$f" = "$unit" ]; then objs+=("$ROOT/tune/obj/${unit}_$name.o"); else objs+=("$SRC/build/$f.o"); fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o "$ROOT/tune/libsphb_$name.so" "${objs[@]}"
grep -E "registers" "$ROOT/tune/obj/${unit}_$name.log" | tr '\n' ' '; echo; echo "built tune/libsphb_$name.so"
