#!/usr/bin/env bash
# Builds the host shell (libsph_host.so: reference-compatible SPHEngine over the C ABI) and the
# pybind11 module `sph` (python/sph.*.so).  Pure g++; links libsphb.so through its C ABI only.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="$(cd "$HERE/.." && pwd)"
ROOT="$(cd "$PKG/.." && pwd)"
CXX=/usr/bin/g++; [ -x "$CXX" ] || CXX=g++
PY="${PYTHON:-python}"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
PBINC="$($PY -c 'import pybind11; print(pybind11.get_include())')"
EXT="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
COMMON=(-std=c++17 -O2 -fPIC -Wall -I"$ROOT/include" -I"$ROOT/include/compat" -I"$HERE")
[ -f "$PKG/libsphb.so" ] || { echo "libsphb.so missing: run csrc/build.sh first" >&2; exit 1; }
"$CXX" "${COMMON[@]}" -shared -o "$PKG/libsph_host.so" "$HERE/sph_host.cpp" -L"$PKG" -lsphb -Wl,-rpath,'$ORIGIN'
"$CXX" "${COMMON[@]}" -fvisibility=hidden -I"$PYINC" -I"$PBINC" -shared -o "$PKG/python/sph$EXT" "$PKG/python/bindings.cpp" \
    -L"$PKG" -lsph_host -lsphb -Wl,-rpath,'$ORIGIN/..'
echo "built $PKG/libsph_host.so and $PKG/python/sph$EXT"
