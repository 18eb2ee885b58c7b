"""CPU: libsphb.so loads and exports every symbol include/sphb.h declares; no compute without a GPU."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "sphb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sphb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(pkg):
    lib = pkg.capi.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"libsphb.so does not export {name}"
    assert sorted(pkg.capi.ABI_SYMBOLS) == names, "capi.ABI_SYMBOLS out of sync with include/sphb.h"
    assert lib.sphb_version() == 100


def test_params_struct_layout(pkg):
    # POD mirror of sph::SPHParameters: 16 floats in the reference's field order (sph_engine.h:14-34)
    assert ctypes.sizeof(pkg.capi.SphbParams) == 64
    assert [f for f, _ in pkg.capi.SphbParams._fields_] == list(pkg.PARAM_FIELDS)
    assert pkg.DEFAULT_PARAMS["neighbor_search_radius"] == 0.04 and pkg.DEFAULT_PARAMS["damping"] == 0.99


def test_stats_struct_layout(pkg):
    # sphb_stats of include/sphb.h: five doubles, then five 64-bit counters, in this order
    assert ctypes.sizeof(pkg.capi.SphbStats) == 80
    text = (ROOT / "include" / "sphb.h").read_text()
    body = re.sub(r"/\*.*?\*/", "", text[text.index("typedef struct sphb_stats"):text.index("} sphb_stats;")], flags=re.S)
    assert re.findall(r"(?:double|uint64_t)\s+(\w+);", body) == [f for f, _ in pkg.capi.SphbStats._fields_]


def test_fails_loudly_without_gpu(pkg, has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.SphbError) as ei:
        pkg.Context(1000, 0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_null_arguments(pkg):
    lib = pkg.capi.load_library()
    assert lib.sphb_create(None, 10, 0) == -1
    assert lib.sphb_step(None, ctypes.c_float(0.001)) == -1
    assert lib.sphb_set_params(None, None) == -1
    lib.sphb_destroy(None)  # no-op
