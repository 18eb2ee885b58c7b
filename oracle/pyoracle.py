"""TEST INFRASTRUCTURE ONLY — ctypes front-end for the parity oracles.

Three shared objects speak the same flat C interface (``<prefix>_create`` ...):

* ``oracle/_ref/liboracle_strict.so``  prefix ``ref``  — the UNMODIFIED reference engine, IEEE-strict
* ``oracle/_ref/liboracle_fast.so``    prefix ``ref``  — the UNMODIFIED reference, its own -ffast-math flags
* ``oracle/liboracle_port.so``         prefix ``port`` — oracle/sph_oracle.c, our plain-C restatement

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The product (include/sphb.h → libsphb.so) never does.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

PARAM_NAMES = (
    "rest_density", "gas_constant", "viscosity", "smoothing_length", "particle_mass",
    "timestep", "gravity", "damping", "CFL_factor",
    "xmin", "xmax", "ymin", "ymax", "zmin", "zmax", "neighbor_search_radius",
)

_fp = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_dp = C.POINTER(C.c_double)


def _f(a):
    return None if a is None else a.ctypes.data_as(_fp)


def lib_path(kind: str) -> Path:
    if kind == "strict":
        return HERE / "_ref" / "liboracle_strict.so"
    if kind == "fast":
        return HERE / "_ref" / "liboracle_fast.so"
    if kind == "port":
        return HERE / "liboracle_port.so"
    raise ValueError(kind)


def available(kind: str) -> bool:
    return lib_path(kind).exists()


_LIBS: dict[str, "OracleLib"] = {}


class OracleLib:
    def __init__(self, kind: str):
        self.kind = kind
        self.prefix = "port" if kind == "port" else "ref"
        self.lib = C.CDLL(str(lib_path(kind)))
        L, p = self.lib, self.prefix

        def sig(name, res, *args):
            fn = getattr(L, f"{p}_{name}")
            fn.restype = res
            fn.argtypes = list(args)
            setattr(self, name, fn)

        vp, sz, f32 = C.c_void_p, C.c_size_t, C.c_float
        sig("create", vp, sz)
        sig("destroy", None, vp)
        sig("sizeof_particle", C.c_int)
        sig("num_threads", C.c_int)
        sig("default_params", None, _fp)
        sig("initialize", None, vp, _fp)
        sig("set_parameters", None, vp, _fp)
        sig("get_parameters", None, vp, _fp)
        sig("set_smoothing_length", None, vp, f32)
        sig("set_gravity", None, vp, f32)
        sig("set_viscosity", None, vp, f32)
        sig("set_boundaries", None, vp, f32, f32, f32, f32, f32, f32)
        sig("initialize_dam_break", None, vp)
        sig("initialize_fluid_drop", None, vp)
        sig("initialize_granular_flow", None, vp)
        sig("clear_particles", None, vp)
        sig("add_particles", None, vp, sz, _fp, _fp, _fp)
        sig("gen_fluid_block", sz, _fp, _fp, f32, f32, sz, _fp, _fp)
        sig("gen_boundary_box", sz, _fp, _fp, f32, f32, sz, _fp, _fp)
        sig("gen_fluid_drop", sz, _fp, f32, f32, f32, sz, _fp, _fp)
        sig("gen_dam_break", sz, _fp, _fp, f32, f32, sz, _fp, _fp)
        sig("size", sz, vp)
        sig("capacity", sz, vp)
        sig("time", f32, vp)
        sig("step_count", sz, vp)
        sig("is_initialized", C.c_int, vp)
        sig("step", None, vp, f32)
        sig("run_steps", None, vp, sz, C.c_int)
        sig("timed_steps", C.c_double, vp, sz, f32)
        sig("cfl_timestep", f32, vp)
        sig("get_state", None, vp, _fp, _fp, _fp, _fp, _fp, _fp)
        sig("set_state", None, vp, _fp, _fp)
        sig("get_keys", None, vp, _u64p)
        sig("get_neighbor_counts", None, vp, _u32p)
        sig("get_neighbor_list", sz, vp, sz, sz, _u32p)
        sig("update_neighbor_lists", None, vp)
        sig("hash_total_cells", sz, vp)
        sig("hash_max_per_cell", sz, vp)
        sig("total_mass", f32, vp)
        sig("total_energy", f32, vp)
        sig("conservation_errors", None, vp, _fp, _fp)
        sig("get_stats", None, vp, _dp)
        sig("reset_stats", None, vp)
        sig("get_densities_raw", sz, vp, sz, _fp)
        sig("set_kernel", C.c_int, vp, C.c_int)
        sig("kernel_W", f32, vp, f32, f32, f32)
        sig("kernel_gradW", None, vp, f32, f32, f32, _fp)
        sig("kernel_lapW", f32, vp, f32, f32, f32)


def load(kind: str) -> OracleLib:
    if kind not in _LIBS:
        _LIBS[kind] = OracleLib(kind)
    return _LIBS[kind]


def default_params(kind: str = "port") -> dict:
    buf = np.zeros(16, np.float32)
    load(kind).default_params(_f(buf))
    return {k: np.float32(v) for k, v in zip(PARAM_NAMES, buf)}


def pack_params(p: dict) -> np.ndarray:
    return np.array([p[k] for k in PARAM_NAMES], dtype=np.float32)


class Engine:
    """One oracle engine instance; mirrors sph::SPHEngine (reference src/sph_engine.h:89-141)."""

    def __init__(self, kind: str = "strict", max_particles: int = 1_000_000):
        self.L = load(kind)
        self.kind = kind
        self.h = self.L.create(max_particles)

    def close(self):
        if self.h:
            self.L.destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- configuration -------------------------------------------------------------------------
    def initialize(self, params: dict | None = None):
        p = dict(default_params(self.kind))
        if params:
            p.update(params)
        self.L.initialize(self.h, _f(pack_params(p)))

    def set_parameters(self, params: dict):
        self.L.set_parameters(self.h, _f(pack_params(params)))

    def get_parameters(self) -> dict:
        buf = np.zeros(16, np.float32)
        self.L.get_parameters(self.h, _f(buf))
        return {k: np.float32(v) for k, v in zip(PARAM_NAMES, buf)}

    def set_smoothing_length(self, h):
        self.L.set_smoothing_length(self.h, float(h))

    def set_kernel(self, kernel_type: int):
        """0 cubic spline (the engine's own), 1 Wendland C2, 2 Gaussian: create_kernel (reference kernels.cpp:224-236)
        installed into the engine's kernel slot; initialize() / set_smoothing_length() put the cubic spline back."""
        if self.L.set_kernel(self.h, int(kernel_type)) != 0:
            raise ValueError(f"unknown kernel type {kernel_type}")

    def set_gravity(self, g):
        self.L.set_gravity(self.h, float(g))

    def set_viscosity(self, mu):
        self.L.set_viscosity(self.h, float(mu))

    def set_boundaries(self, x0, x1, y0, y1, z0, z1):
        self.L.set_boundaries(self.h, float(x0), float(x1), float(y0), float(y1), float(z0), float(z1))

    def initialize_dam_break(self):
        self.L.initialize_dam_break(self.h)

    def initialize_fluid_drop(self):
        self.L.initialize_fluid_drop(self.h)

    def initialize_granular_flow(self):
        self.L.initialize_granular_flow(self.h)

    def clear_particles(self):
        self.L.clear_particles(self.h)

    def add_particles(self, pos, vel=None, mass=None):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        n = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32).reshape(-1, 3)
        mass = None if mass is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mass, np.float32), (n,)))
        self.L.add_particles(self.h, n, _f(pos), _f(vel), _f(mass))

    # --- stepping ------------------------------------------------------------------------------
    def step(self, dt: float = 0.0):
        self.L.step(self.h, float(dt))

    def run_steps(self, n, adaptive=True):
        self.L.run_steps(self.h, int(n), int(bool(adaptive)))

    def timed_steps(self, n, dt) -> float:
        return float(self.L.timed_steps(self.h, int(n), float(dt)))

    def cfl_timestep(self) -> float:
        return float(self.L.cfl_timestep(self.h))

    def update_neighbor_lists(self):
        self.L.update_neighbor_lists(self.h)

    # --- observation ---------------------------------------------------------------------------
    @property
    def size(self) -> int:
        return int(self.L.size(self.h))

    @property
    def capacity(self) -> int:
        return int(self.L.capacity(self.h))

    @property
    def time(self) -> float:
        return float(self.L.time(self.h))

    @property
    def step_count(self) -> int:
        return int(self.L.step_count(self.h))

    def state(self) -> dict:
        n = self.size
        out = {
            "pos": np.zeros((n, 3), np.float32), "vel": np.zeros((n, 3), np.float32),
            "mass": np.zeros(n, np.float32), "rho": np.zeros(n, np.float32),
            "P": np.zeros(n, np.float32), "acc": np.zeros((n, 3), np.float32),
        }
        self.L.get_state(self.h, _f(out["pos"]), _f(out["vel"]), _f(out["mass"]), _f(out["rho"]), _f(out["P"]), _f(out["acc"]))
        return out

    def set_state(self, pos=None, vel=None):
        pos = None if pos is None else np.ascontiguousarray(pos, np.float32)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        self.L.set_state(self.h, _f(pos), _f(vel))

    def keys(self) -> np.ndarray:
        k = np.zeros(self.size, np.uint64)
        self.L.get_keys(self.h, k.ctypes.data_as(_u64p))
        return k

    def neighbor_counts(self) -> np.ndarray:
        c = np.zeros(self.size, np.uint32)
        self.L.get_neighbor_counts(self.h, c.ctypes.data_as(_u32p))
        return c

    def neighbor_list(self, i: int) -> np.ndarray:
        cap = 1 << 16
        buf = np.zeros(cap, np.uint32)
        n = int(self.L.get_neighbor_list(self.h, i, cap, buf.ctypes.data_as(_u32p)))
        return buf[: min(n, cap)].copy()

    def stats(self) -> dict:
        s = np.zeros(7, np.float64)
        self.L.get_stats(self.h, s.ctypes.data_as(_dp))
        return dict(total_time=s[0], neighbor_search_time=s[1], density_computation_time=s[2],
                    force_computation_time=s[3], integration_time=s[4], max_neighbors=int(s[5]),
                    total_neighbor_queries=int(s[6]))

    def reset_stats(self):
        self.L.reset_stats(self.h)

    def total_mass(self) -> float:
        return float(self.L.total_mass(self.h))

    def total_energy(self) -> float:
        return float(self.L.total_energy(self.h))

    def conservation_errors(self):
        a, b = C.c_float(), C.c_float()
        self.L.conservation_errors(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def densities_raw(self) -> np.ndarray:
        cap = self.capacity
        buf = np.zeros(cap, np.float32)
        n = int(self.L.get_densities_raw(self.h, cap, _f(buf)))
        return buf[:n]

    def kernel_W(self, r):
        return float(self.L.kernel_W(self.h, *map(float, r)))

    def kernel_gradW(self, r):
        out = np.zeros(3, np.float32)
        self.L.kernel_gradW(self.h, *map(float, r), _f(out))
        return out

    def kernel_lapW(self, r):
        return float(self.L.kernel_lapW(self.h, *map(float, r)))


# --- lattice generators (reference particle.cpp:166-229, sph_engine.cpp:450-514) ----------------
def _gen(kind, fn_name, cap_guess, *args):
    L = load(kind)
    fn = getattr(L, fn_name)
    cap = int(cap_guess)
    while True:
        pos = np.zeros((cap, 3), np.float32)
        mass = np.zeros(cap, np.float32)
        n = int(fn(*args, cap, _f(pos), _f(mass)))
        if n <= cap:
            return pos[:n].copy(), mass[:n].copy()
        cap = n


def gen_fluid_block(center, size, spacing, mass=1.0, kind="port"):
    c = np.asarray(center, np.float32); s = np.asarray(size, np.float32)
    guess = int(np.prod(np.floor(s / np.float32(spacing)) + 1)) + 16
    return _gen(kind, "gen_fluid_block", guess, _f(c), _f(s), float(spacing), float(mass))


def gen_boundary_box(center, size, spacing, mass=1.0, kind="port"):
    c = np.asarray(center, np.float32); s = np.asarray(size, np.float32)
    n = np.floor(s / np.float32(spacing)) + 1
    guess = int(2 * (n[0] * n[1] + n[1] * n[2] + n[0] * n[2])) + 16
    return _gen(kind, "gen_boundary_box", guess, _f(c), _f(s), float(spacing), float(mass))


def gen_fluid_drop(center, radius, spacing, mass=1.0, kind="port"):
    c = np.asarray(center, np.float32)
    guess = int((2 * radius / spacing + 1) ** 3) + 16
    return _gen(kind, "gen_fluid_drop", guess, _f(c), float(radius), float(spacing), float(mass))


def gen_dam_break(dam, fluid, spacing, mass=1.0, kind="port"):
    d = np.asarray(dam, np.float32); f = np.asarray(fluid, np.float32)
    nb = np.floor(d / np.float32(spacing)) + 1
    nf = np.floor(f / np.float32(spacing)) + 1
    guess = int(2 * (nb[0] * nb[1] + nb[1] * nb[2] + nb[0] * nb[2]) + np.prod(nf)) + 16
    return _gen(kind, "gen_dam_break", guess, _f(d), _f(f), float(spacing), float(mass))


def set_threads(n: int):
    os.environ["OMP_NUM_THREADS"] = str(n)
