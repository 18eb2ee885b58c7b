#!/usr/bin/env bash
# Drop-in proof: compile the reference's OWN benchmark driver (benchmarks/performance_test.cpp), unmodified and
# from where it lies, against the B200 host shell instead of the reference engine.  Output: dropin/_ref/ (git-ignored,
# travels to the GPU box like oracle/_ref).  Needs /root/reference; a no-op where it is absent.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/.." && pwd)"
REF="${REF:-/root/reference}"
PKG="$ROOT/sph-particle-simulator_b200"
CXX=/usr/bin/g++; [ -x "$CXX" ] || CXX=g++
if [ ! -f "$REF/benchmarks/performance_test.cpp" ]; then echo "dropin: $REF not present — keeping prebuilt dropin/_ref if any"; exit 0; fi
mkdir -p "$HERE/_ref"
"$CXX" -std=c++17 -O2 -I"$PKG/host/fwd" -I"$PKG/host" -I"$ROOT/include" -I"$ROOT/include/compat" \
    "$REF/benchmarks/performance_test.cpp" -o "$HERE/_ref/performance_test" \
    -L"$PKG" -lsph_host -lsphb -Wl,-rpath,"\$ORIGIN/../../sph-particle-simulator_b200"
echo "built $HERE/_ref/performance_test (reference benchmarks/performance_test.cpp on the B200 engine)"
