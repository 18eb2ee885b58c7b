"""B200-native SPH hot path behind the reference's SPHEngine / sph.Simulator API.

The directory name carries a hyphen (it mirrors the upstream project name), so it is loaded through
``__graft_entry__.load_package()`` under the module name ``sph_b200`` rather than by a plain import.

Contents: ``csrc/`` (CUDA kernels + the C ABI of include/sphb.h → libsphb.so), ``capi`` (ctypes binding
of that ABI), ``host/`` (C++ SPHEngine shell with the reference's class surface) and ``python/`` (pybind11
module ``sph`` with the reference's Python surface).
"""
from . import capi  # noqa: F401
from .capi import Context, MultiContext, SphbError, DEFAULT_PARAMS, PARAM_FIELDS  # noqa: F401
