"""ctypes binding of the C ABI in include/sphb.h (libsphb.so).

Thin by design: one Python method per C entry point, numpy arrays in and out, every non-zero
return code raised as :class:`SphbError` with the library's own message.  There is no fallback of
any kind — if the shared library is missing or no CUDA device is usable this module raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libsphb.so"

PARAM_FIELDS = (
    "rest_density", "gas_constant", "viscosity", "smoothing_length", "particle_mass",
    "timestep", "gravity", "damping", "CFL_factor",
    "xmin", "xmax", "ymin", "ymax", "zmin", "zmax", "neighbor_search_radius",
)

# reference defaults, src/sph_engine.h:14-34
DEFAULT_PARAMS = dict(
    rest_density=1000.0, gas_constant=2000.0, viscosity=0.001, smoothing_length=0.02,
    particle_mass=0.001, timestep=0.001, gravity=-9.81, damping=0.99, CFL_factor=0.4,
    xmin=-1.0, xmax=1.0, ymin=-1.0, ymax=1.0, zmin=-1.0, zmax=1.0, neighbor_search_radius=0.04,
)

OPT_MATH_MODE = 1
OPT_WALK_RADIUS = 2
OPT_STAGE_TIMING = 3
OPT_DEBUG_CAPTURE = 4
OPT_PAIR_KERNEL = 5
OPT_GRID_REFINE = 6
OPT_LAYOUT_MAJOR = 7
OPT_PAIR_MODE = 8
OPT_KERNEL_TYPE = 9
OPT_STEP_GRAPHS = 10
OPT_LANES_PER_PARTICLE = 11
KERNEL_CUBIC_SPLINE, KERNEL_WENDLAND_C2, KERNEL_GAUSSIAN = 0, 1, 2
OPT_MULTI_AXIS = 100
OPT_MULTI_HALO_LAYERS = 101
OPT_MULTI_REBALANCE_MIN = 102
MATH_STRICT = 0
MATH_FAST = 1

# every symbol include/sphb.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = (
    "sphb_create", "sphb_destroy", "sphb_last_error", "sphb_version", "sphb_set_option", "sphb_get_option",
    "sphb_set_stream", "sphb_synchronize", "sphb_set_params", "sphb_get_params", "sphb_upload",
    "sphb_upload_strided", "sphb_download", "sphb_download_begin", "sphb_download_end", "sphb_download_strided", "sphb_size", "sphb_step", "sphb_run_steps",
    "sphb_get_time", "sphb_set_time", "sphb_cfl_timestep", "sphb_get_stats", "sphb_reset_stats",
    "sphb_diagnostics", "sphb_set_colors", "sphb_export_instances", "sphb_debug_dump", "sphb_debug_stencil",
    "sphb_set_slab", "sphb_upload_ids", "sphb_slab_append", "sphb_slab_exchange_pack", "sphb_slab_exchange_count",
    "sphb_slab_exchange_split", "sphb_get_cfl_state", "sphb_set_cfl_state",
    "sphb_slab_download", "sphb_slab_download_begin", "sphb_slab_download_end", "sphb_read_small",
    "sphb_create_multi", "sphb_destroy_multi", "sphb_multi_last_error", "sphb_multi_device_count", "sphb_multi_set_option",
    "sphb_multi_set_params", "sphb_multi_upload", "sphb_multi_upload_strided", "sphb_multi_step", "sphb_multi_run_steps",
    "sphb_multi_synchronize", "sphb_multi_size", "sphb_multi_download", "sphb_multi_download_strided", "sphb_multi_get_time",
    "sphb_multi_set_time", "sphb_multi_cfl_timestep", "sphb_multi_get_stats", "sphb_multi_reset_stats", "sphb_multi_diagnostics",
    "sphb_multi_layout",
)


class SphbParams(C.Structure):
    _fields_ = [(k, C.c_float) for k in PARAM_FIELDS]


class SphbStats(C.Structure):
    _fields_ = [
        ("total_time", C.c_double), ("neighbor_search_time", C.c_double),
        ("density_computation_time", C.c_double), ("force_computation_time", C.c_double),
        ("integration_time", C.c_double), ("max_neighbors", C.c_uint64),
        ("total_neighbor_queries", C.c_uint64), ("steps", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("error_flags", C.c_uint64),
    ]


class SphbSlab(C.Structure):
    _fields_ = [("axis", C.c_int32), ("own_lo", C.c_int32), ("own_hi", C.c_int32), ("halo_layers", C.c_int32),
                ("id_space", C.c_uint64), ("box_min", C.c_float * 3), ("box_max", C.c_float * 3)]


class SphbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sphb error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """Load libsphb.so (built in-tree by csrc/build.sh); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = Path(os.environ.get("SPHB_LIB", str(LIB_PATH)))   # SPHB_LIB: alternative build of the same library (tuning runs)
    if not path.exists():
        raise FileNotFoundError(f"{LIB_PATH} not built — run __graft_entry__.build() (sph-particle-simulator_b200/csrc/build.sh)")
    L = C.CDLL(str(path))
    vp, sz, fp = C.c_void_p, C.c_size_t, C.POINTER(C.c_float)
    L.sphb_version.restype = C.c_int
    L.sphb_last_error.restype = C.c_char_p
    L.sphb_last_error.argtypes = [vp]
    L.sphb_create.argtypes = [C.POINTER(vp), sz, C.c_int]
    L.sphb_destroy.argtypes = [vp]
    L.sphb_destroy.restype = None
    L.sphb_set_option.argtypes = [vp, C.c_int, C.c_int64]
    L.sphb_get_option.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    L.sphb_set_stream.argtypes = [vp, vp]
    L.sphb_synchronize.argtypes = [vp]
    L.sphb_set_params.argtypes = [vp, C.POINTER(SphbParams)]
    L.sphb_get_params.argtypes = [vp, C.POINTER(SphbParams)]
    L.sphb_upload.argtypes = [vp, sz, vp, vp, vp]
    L.sphb_upload_strided.argtypes = [vp, sz, vp, sz, sz, sz, sz]
    L.sphb_download.argtypes = [vp, vp, vp, vp, vp, vp]
    L.sphb_download_begin.argtypes = [vp, vp, vp, vp, vp, vp]
    L.sphb_download_end.argtypes = [vp]
    L.sphb_download_strided.argtypes = [vp, vp, sz, sz, sz, sz, sz]
    L.sphb_size.argtypes = [vp, C.POINTER(sz)]
    L.sphb_step.argtypes = [vp, C.c_float]
    L.sphb_run_steps.argtypes = [vp, sz, C.c_float]
    L.sphb_get_time.argtypes = [vp, fp, C.POINTER(C.c_uint64)]
    L.sphb_set_time.argtypes = [vp, C.c_float, C.c_uint64]
    L.sphb_cfl_timestep.argtypes = [vp, fp]
    L.sphb_get_stats.argtypes = [vp, C.POINTER(SphbStats)]
    L.sphb_reset_stats.argtypes = [vp]
    L.sphb_diagnostics.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), fp]
    L.sphb_set_colors.argtypes = [vp, sz, vp]
    L.sphb_export_instances.argtypes = [vp, vp, C.c_int, vp]
    L.sphb_debug_dump.argtypes = [vp, vp, vp, vp]
    L.sphb_debug_stencil.argtypes = [C.c_int, vp, C.POINTER(C.c_float)]
    L.sphb_set_slab.argtypes = [vp, C.POINTER(SphbSlab)]
    L.sphb_upload_ids.argtypes = [vp, sz, vp, vp, vp, vp]
    L.sphb_get_cfl_state.argtypes = [vp, fp, fp, C.POINTER(C.c_int)]
    L.sphb_set_cfl_state.argtypes = [vp, C.c_float, fp]
    L.sphb_slab_exchange_pack.argtypes = [vp, vp, C.c_int, C.c_int, vp, sz, vp]
    L.sphb_slab_exchange_count.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.sphb_slab_exchange_split.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, sz]
    L.sphb_slab_append.argtypes = [vp, vp, sz, C.c_int]
    L.sphb_slab_download.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, C.POINTER(sz)]
    L.sphb_slab_download_begin.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp]
    L.sphb_slab_download_end.argtypes = [vp, C.POINTER(sz)]
    L.sphb_read_small.argtypes = [vp, vp, vp, sz]
    L.sphb_create_multi.argtypes = [C.POINTER(vp), sz, C.c_int, C.POINTER(C.c_int)]
    L.sphb_destroy_multi.argtypes = [vp]
    L.sphb_destroy_multi.restype = None
    L.sphb_multi_last_error.restype = C.c_char_p
    L.sphb_multi_last_error.argtypes = [vp]
    L.sphb_multi_device_count.argtypes = [vp]
    L.sphb_multi_set_option.argtypes = [vp, C.c_int, C.c_int64]
    L.sphb_multi_set_params.argtypes = [vp, C.POINTER(SphbParams)]
    L.sphb_multi_upload.argtypes = [vp, sz, vp, vp, vp]
    L.sphb_multi_upload_strided.argtypes = [vp, sz, vp, sz, sz, sz, sz]
    L.sphb_multi_step.argtypes = [vp, C.c_float]
    L.sphb_multi_run_steps.argtypes = [vp, sz, C.c_float]
    L.sphb_multi_synchronize.argtypes = [vp]
    L.sphb_multi_size.argtypes = [vp, C.POINTER(sz)]
    L.sphb_multi_download.argtypes = [vp, vp, vp, vp, vp, vp]
    L.sphb_multi_download_strided.argtypes = [vp, vp, sz, sz, sz, sz, sz]
    L.sphb_multi_get_time.argtypes = [vp, fp, C.POINTER(C.c_uint64)]
    L.sphb_multi_set_time.argtypes = [vp, C.c_float, C.c_uint64]
    L.sphb_multi_cfl_timestep.argtypes = [vp, fp]
    L.sphb_multi_get_stats.argtypes = [vp, C.POINTER(SphbStats)]
    L.sphb_multi_reset_stats.argtypes = [vp]
    L.sphb_multi_diagnostics.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), fp]
    L.sphb_multi_layout.argtypes = [vp, vp, C.POINTER(C.c_int), vp, vp]
    _lib = L
    return L


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def debug_stencil(radius: int):
    """(reach table int8[(2R+1), (2R+1)], cell scale) of the fast path's static spherical stencil — no GPU needed."""
    L = load_library()
    R = int(radius)
    reach = np.zeros((2 * R + 1) ** 2, np.int8)
    scale = C.c_float()
    n = L.sphb_debug_stencil(R, _ptr(reach), C.byref(scale))
    if n < 0:
        raise ValueError(f"no stencil for radius {R}")
    return reach.reshape(2 * R + 1, 2 * R + 1), float(scale.value)


class Context:
    """One sphb_ctx: the device-side state of one engine on one GPU."""

    def __init__(self, capacity: int, device: int = 0):
        self.L = load_library()
        self.capacity = int(capacity)
        h = C.c_void_p()
        rc = self.L.sphb_create(C.byref(h), self.capacity, int(device))
        if rc != 0:
            raise SphbError(rc, (self.L.sphb_last_error(None) or b"").decode())
        self.h = h

    def _ck(self, rc: int):
        if rc != 0:
            raise SphbError(rc, (self.L.sphb_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.sphb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- options / configuration ----------------------------------------------------------------
    def set_option(self, opt: int, value: int):
        self._ck(self.L.sphb_set_option(self.h, opt, int(value)))

    def get_option(self, opt: int) -> int:
        v = C.c_int64()
        self._ck(self.L.sphb_get_option(self.h, opt, C.byref(v)))
        return v.value

    def set_stream(self, cuda_stream: int):
        self._ck(self.L.sphb_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def synchronize(self):
        self._ck(self.L.sphb_synchronize(self.h))

    def set_params(self, params: dict):
        p = SphbParams(**{k: float(params[k]) for k in PARAM_FIELDS})
        self._ck(self.L.sphb_set_params(self.h, C.byref(p)))

    def get_params(self) -> dict:
        p = SphbParams()
        self._ck(self.L.sphb_get_params(self.h, C.byref(p)))
        return {k: np.float32(getattr(p, k)) for k in PARAM_FIELDS}

    # ---- state ------------------------------------------------------------------------------------
    def upload(self, pos, vel=None, mass=None):
        """pos (n,3), vel (n,3) or None, mass (n,) or None — numpy float32 (copied if not contiguous)."""
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        n = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32).reshape(n, 3)
        mass = None if mass is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mass, np.float32), (n,)))
        self._ck(self.L.sphb_upload(self.h, n, _ptr(pos), _ptr(vel), _ptr(mass)))

    def upload_raw(self, n: int, pos_ptr: int, vel_ptr: int | None, mass_ptr: int | None):
        """Upload from raw host pointers (e.g. pinned torch tensors' data_ptr())."""
        self._ck(self.L.sphb_upload(self.h, int(n), _ptr(pos_ptr), _ptr(vel_ptr), _ptr(mass_ptr)))

    def upload_strided(self, n: int, base_ptr: int, stride: int, off_pos: int, off_vel: int, off_mass: int):
        self._ck(self.L.sphb_upload_strided(self.h, int(n), C.c_void_p(base_ptr), stride, off_pos, off_vel, off_mass))

    @property
    def size(self) -> int:
        n = C.c_size_t()
        self._ck(self.L.sphb_size(self.h, C.byref(n)))
        return n.value

    def download(self, pos=True, vel=True, rho=True, pressure=True, acc=True) -> dict:
        n = self.size
        out = {}
        if pos: out["pos"] = np.zeros((n, 3), np.float32)
        if vel: out["vel"] = np.zeros((n, 3), np.float32)
        if rho: out["rho"] = np.zeros(n, np.float32)
        if pressure: out["P"] = np.zeros(n, np.float32)
        if acc: out["acc"] = np.zeros((n, 3), np.float32)
        self._ck(self.L.sphb_download(self.h, _ptr(out.get("pos")), _ptr(out.get("vel")), _ptr(out.get("rho")),
                                      _ptr(out.get("P")), _ptr(out.get("acc"))))
        return out

    def download_raw(self, pos_ptr=None, vel_ptr=None, rho_ptr=None, p_ptr=None, acc_ptr=None):
        self._ck(self.L.sphb_download(self.h, _ptr(pos_ptr), _ptr(vel_ptr), _ptr(rho_ptr), _ptr(p_ptr), _ptr(acc_ptr)))

    def download_begin_raw(self, pos_ptr=None, vel_ptr=None, rho_ptr=None, p_ptr=None, acc_ptr=None):
        """Starts the download into (pinned) host memory and returns; download_end() waits for it."""
        self._ck(self.L.sphb_download_begin(self.h, _ptr(pos_ptr), _ptr(vel_ptr), _ptr(rho_ptr), _ptr(p_ptr), _ptr(acc_ptr)))

    def download_end(self):
        self._ck(self.L.sphb_download_end(self.h))

    def download_strided(self, base_ptr: int, stride: int, off_pos, off_vel, off_density, off_pressure):
        none = (1 << 64) - 1
        f = lambda o: none if o is None else int(o)
        self._ck(self.L.sphb_download_strided(self.h, C.c_void_p(base_ptr), stride, f(off_pos), f(off_vel), f(off_density),
                                              f(off_pressure)))

    # ---- stepping -----------------------------------------------------------------------------------
    def step(self, dt: float = 0.0):
        self._ck(self.L.sphb_step(self.h, float(dt)))

    def run_steps(self, n: int, dt: float = 0.0):
        self._ck(self.L.sphb_run_steps(self.h, int(n), float(dt)))

    def get_time(self):
        t, s = C.c_float(), C.c_uint64()
        self._ck(self.L.sphb_get_time(self.h, C.byref(t), C.byref(s)))
        return t.value, s.value

    def set_time(self, t: float, step_count: int):
        self._ck(self.L.sphb_set_time(self.h, float(t), int(step_count)))

    def cfl_timestep(self) -> float:
        dt = C.c_float()
        self._ck(self.L.sphb_cfl_timestep(self.h, C.byref(dt)))
        return dt.value

    # ---- diagnostics ----------------------------------------------------------------------------
    def stats(self) -> dict:
        s = SphbStats()
        self._ck(self.L.sphb_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in SphbStats._fields_}

    def reset_stats(self):
        self._ck(self.L.sphb_reset_stats(self.h))

    def diagnostics(self):
        a, b, c = C.c_double(), C.c_double(), C.c_float()
        self._ck(self.L.sphb_diagnostics(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set_colors(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
        self._ck(self.L.sphb_set_colors(self.h, rgb.shape[0], _ptr(rgb)))

    def export_instances(self, default_rgb=None, device_ptr: int | None = None):
        """The renderer's instance records (n, 9): position, velocity, colour.  device_ptr: write into device memory instead."""
        d = None if default_rgb is None else np.ascontiguousarray(default_rgb, np.float32)
        if device_ptr is not None:
            self._ck(self.L.sphb_export_instances(self.h, C.c_void_p(device_ptr), 1, _ptr(d)))
            return None
        out = np.zeros((self.size, 9), np.float32)
        self._ck(self.L.sphb_export_instances(self.h, _ptr(out), 0, _ptr(d)))
        return out

    def debug_dump(self, keys=True, perm=True, counts=True) -> dict:
        n = self.size
        k = np.zeros(n, np.uint64) if keys else None
        p = np.zeros(n, np.uint32) if perm else None
        c = np.zeros(n, np.uint32) if counts else None
        self._ck(self.L.sphb_debug_dump(self.h, _ptr(k), _ptr(p), _ptr(c)))
        return {"keys": k, "perm": p, "nbr_count": c}

    # ---- slab decomposition (multi-GPU) -----------------------------------------------------------
    def set_slab(self, axis: int, own_lo: int, own_hi: int, halo_layers: int, id_space: int, box_min, box_max):
        sl = SphbSlab(int(axis), int(own_lo), int(own_hi), int(halo_layers), int(id_space),
                      (C.c_float * 3)(*map(float, box_min)), (C.c_float * 3)(*map(float, box_max)))
        self._ck(self.L.sphb_set_slab(self.h, C.byref(sl)))

    def clear_slab(self):
        self._ck(self.L.sphb_set_slab(self.h, None))

    def upload_ids(self, pos, vel, mass, ids):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        n = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32).reshape(n, 3)
        mass = None if mass is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mass, np.float32), (n,)))
        ids = np.ascontiguousarray(ids, np.uint32).reshape(n)
        self._ck(self.L.sphb_upload_ids(self.h, n, _ptr(pos), _ptr(vel), _ptr(mass), _ptr(ids)))

    def slab_exchange_pack(self, cuts, my_rank: int, d_out_ptr: int, cap_records: int) -> np.ndarray:
        cuts = np.ascontiguousarray(cuts, np.int32)
        nranks = cuts.shape[0] - 1
        counts = np.zeros(2 * nranks, np.uint64)
        self._ck(self.L.sphb_slab_exchange_pack(self.h, _ptr(cuts), nranks, int(my_rank), C.c_void_p(d_out_ptr), int(cap_records),
                                                _ptr(counts)))
        return counts

    def slab_exchange_count(self, cuts, my_rank: int, d_counts_ptr: int):
        """Phase 1 (asynchronous): this rank's 2*nranks group sizes (uint32) land in the device buffer d_counts_ptr."""
        cuts = np.ascontiguousarray(cuts, np.int32)
        self._ck(self.L.sphb_slab_exchange_count(self.h, _ptr(cuts), cuts.shape[0] - 1, int(my_rank), C.c_void_p(d_counts_ptr)))

    def slab_exchange_split(self, cuts, my_rank: int, counts, d_out_ptr: int, cap_records: int):
        """Phase 2: move the records, given this rank's group sizes (host array, as gathered from phase 1)."""
        cuts = np.ascontiguousarray(cuts, np.int32)
        counts = np.ascontiguousarray(counts, np.uint32)
        self._ck(self.L.sphb_slab_exchange_split(self.h, _ptr(cuts), cuts.shape[0] - 1, int(my_rank), _ptr(counts),
                                                 C.c_void_p(d_out_ptr), int(cap_records)))

    def slab_append(self, d_in_ptr: int, count: int, ghost: bool):
        flag = -1 if ghost is None else (1 if ghost else 0)     # None: records carry their own ghost flag
        self._ck(self.L.sphb_slab_append(self.h, C.c_void_p(d_in_ptr), int(count), flag))

    def slab_download(self, pos=True, vel=True, rho=True, pressure=True, acc=True) -> dict:
        cap = self.size
        out = {"ids": np.zeros(cap, np.uint32)}
        if pos: out["pos"] = np.zeros((cap, 3), np.float32)
        if vel: out["vel"] = np.zeros((cap, 3), np.float32)
        if rho: out["rho"] = np.zeros(cap, np.float32)
        if pressure: out["P"] = np.zeros(cap, np.float32)
        if acc: out["acc"] = np.zeros((cap, 3), np.float32)
        n = C.c_size_t()
        self._ck(self.L.sphb_slab_download(self.h, cap, _ptr(out["ids"]), _ptr(out.get("pos")), _ptr(out.get("vel")),
                                           _ptr(out.get("rho")), _ptr(out.get("P")), _ptr(out.get("acc")), C.byref(n)))
        return {k: v[: n.value] for k, v in out.items()}

    def slab_download_begin_raw(self, cap: int, ids_ptr, pos_ptr=None, vel_ptr=None, rho_ptr=None, p_ptr=None, acc_ptr=None):
        """Starts the read-back of the owned particles into (pinned) host memory; slab_download_end() waits, returns their number."""
        self._ck(self.L.sphb_slab_download_begin(self.h, int(cap), _ptr(ids_ptr), _ptr(pos_ptr), _ptr(vel_ptr), _ptr(rho_ptr), _ptr(p_ptr),
                                                 _ptr(acc_ptr)))

    def read_small(self, d_src_ptr: int, h_dst_pinned_ptr: int, nbytes: int):
        """Enqueue a kernel-written read-back of a small device buffer into pinned host memory (see sphb_read_small)."""
        self._ck(self.L.sphb_read_small(self.h, C.c_void_p(d_src_ptr), C.c_void_p(h_dst_pinned_ptr), int(nbytes)))

    def slab_download_end(self) -> int:
        n = C.c_size_t()
        self._ck(self.L.sphb_slab_download_end(self.h, C.byref(n)))
        return n.value

    def get_cfl_state(self):
        """(max |v|^2 over owned particles, a0[3], a0_fresh) — see sphb_get_cfl_state."""
        v2 = C.c_float()
        a0 = (C.c_float * 3)()
        fresh = C.c_int()
        self._ck(self.L.sphb_get_cfl_state(self.h, C.byref(v2), a0, C.byref(fresh)))
        return np.float32(v2.value), np.array(list(a0), np.float32), bool(fresh.value)

    def set_cfl_state(self, max_v2, a0):
        arr = (C.c_float * 3)(*[float(x) for x in a0])
        self._ck(self.L.sphb_set_cfl_state(self.h, C.c_float(float(max_v2)), arr))


class MultiContext:
    """One sphb_multi: ONE engine over several GPUs of a node (or several slabs on one GPU), driven from this process
    through the C ABI alone (csrc/multi.cu) — no torch.distributed, no NCCL.  Same method names as :class:`Context`."""

    def __init__(self, capacity: int, devices):
        self.L = load_library()
        self.capacity = int(capacity)
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        rc = self.L.sphb_create_multi(C.byref(h), self.capacity, len(self.devices), arr)
        if rc != 0:
            raise SphbError(rc, (self.L.sphb_multi_last_error(None) or b"").decode())
        self.h = h

    def _ck(self, rc: int):
        if rc != 0:
            raise SphbError(rc, (self.L.sphb_multi_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.sphb_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_count(self) -> int:
        return self.L.sphb_multi_device_count(self.h)

    def set_option(self, opt: int, value: int):
        self._ck(self.L.sphb_multi_set_option(self.h, opt, int(value)))

    def set_params(self, params: dict):
        p = SphbParams(**{k: float(params[k]) for k in PARAM_FIELDS})
        self._ck(self.L.sphb_multi_set_params(self.h, C.byref(p)))

    def synchronize(self):
        self._ck(self.L.sphb_multi_synchronize(self.h))

    def upload(self, pos, vel=None, mass=None):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        n = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32).reshape(n, 3)
        mass = None if mass is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mass, np.float32), (n,)))
        self._ck(self.L.sphb_multi_upload(self.h, n, _ptr(pos), _ptr(vel), _ptr(mass)))

    def upload_strided(self, n: int, base_ptr: int, stride: int, off_pos: int, off_vel: int, off_mass: int):
        self._ck(self.L.sphb_multi_upload_strided(self.h, int(n), C.c_void_p(base_ptr), stride, off_pos, off_vel, off_mass))

    @property
    def size(self) -> int:
        n = C.c_size_t()
        self._ck(self.L.sphb_multi_size(self.h, C.byref(n)))
        return n.value

    def download(self, pos=True, vel=True, rho=True, pressure=True, acc=True) -> dict:
        n = self.size
        out = {}
        if pos: out["pos"] = np.zeros((n, 3), np.float32)
        if vel: out["vel"] = np.zeros((n, 3), np.float32)
        if rho: out["rho"] = np.zeros(n, np.float32)
        if pressure: out["P"] = np.zeros(n, np.float32)
        if acc: out["acc"] = np.zeros((n, 3), np.float32)
        self._ck(self.L.sphb_multi_download(self.h, _ptr(out.get("pos")), _ptr(out.get("vel")), _ptr(out.get("rho")),
                                            _ptr(out.get("P")), _ptr(out.get("acc"))))
        return out

    def download_strided(self, base_ptr: int, stride: int, off_pos, off_vel, off_density, off_pressure):
        self._ck(self.L.sphb_multi_download_strided(self.h, C.c_void_p(base_ptr), stride, int(off_pos), int(off_vel),
                                                    int(off_density), int(off_pressure)))

    def step(self, dt: float = 0.0):
        self._ck(self.L.sphb_multi_step(self.h, float(dt)))

    def run_steps(self, n: int, dt: float = 0.0):
        self._ck(self.L.sphb_multi_run_steps(self.h, int(n), float(dt)))

    def get_time(self):
        t, s = C.c_float(), C.c_uint64()
        self._ck(self.L.sphb_multi_get_time(self.h, C.byref(t), C.byref(s)))
        return t.value, s.value

    def set_time(self, t: float, step_count: int):
        self._ck(self.L.sphb_multi_set_time(self.h, float(t), int(step_count)))

    def cfl_timestep(self) -> float:
        dt = C.c_float()
        self._ck(self.L.sphb_multi_cfl_timestep(self.h, C.byref(dt)))
        return dt.value

    def stats(self) -> dict:
        s = SphbStats()
        self._ck(self.L.sphb_multi_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in SphbStats._fields_}

    def reset_stats(self):
        self._ck(self.L.sphb_multi_reset_stats(self.h))

    def diagnostics(self):
        a, b, c = C.c_double(), C.c_double(), C.c_float()
        self._ck(self.L.sphb_multi_diagnostics(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def layout(self) -> dict:
        """Cuts (reference cells on the slab axis), the axis and per-device owned / halo-copy counts of the last exchange."""
        g = self.device_count
        cuts = np.zeros(g + 1, np.int32)
        owned, ghosts = np.zeros(g, np.uint64), np.zeros(g, np.uint64)
        axis = C.c_int()
        self._ck(self.L.sphb_multi_layout(self.h, _ptr(cuts), C.byref(axis), _ptr(owned), _ptr(ghosts)))
        return {"cuts": cuts, "axis": axis.value, "owned": owned, "ghosts": ghosts}
