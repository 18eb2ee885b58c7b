"""Shared helpers of the parity tests."""
from __future__ import annotations

from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

PARAM_NAMES = (
    "rest_density", "gas_constant", "viscosity", "smoothing_length", "particle_mass",
    "timestep", "gravity", "damping", "CFL_factor",
    "xmin", "xmax", "ymin", "ymax", "zmin", "zmax", "neighbor_search_radius",
)

STEP_FIXTURES = (
    "micro_pair", "micro_coincident", "micro_single", "micro_outside", "micro_lattice27", "cloud600",
    "cloud600_truncated_support", "cloud600_wide_cell", "dam_break_13k_tame", "fluid_drop_default_pref",
    "dam_break_example_10k_pref",
)


# the reference engine stepping through its other kernel classes (tests/golden/make_golden.py section 6); "kernel_type" inside
KERNEL_FIXTURES = ("cloud600_wendland", "cloud600_gaussian", "dam_break_13k_wendland", "dam_break_13k_gaussian")


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def params_from(vec) -> dict:
    return {k: np.float32(v) for k, v in zip(PARAM_NAMES, vec)}


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype == np.float32:
        return bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
    return bool(np.array_equal(a, b))


def assert_bits(a, b, what):
    if bits_equal(a, b):
        return
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = np.flatnonzero((a.view(np.uint32) != b.view(np.uint32)).reshape(a.shape[0], -1).any(axis=1)) if a.dtype == np.float32 \
        else np.flatnonzero((a != b).reshape(a.shape[0], -1).any(axis=1))
    i = int(bad[0])
    raise AssertionError(f"{what}: {bad.size}/{a.shape[0]} rows differ bitwise; first at {i}: got {a[i]!r} want {b[i]!r}")


def golden_steps(g: dict):
    """Indices k for which the fixture holds a full 's{k}_*' record."""
    return sorted({int(k[1:].split('_')[0]) for k in g if k.startswith('s') and k[1].isdigit()})


def stable_perm(keys: np.ndarray) -> np.ndarray:
    """ids stably sorted by the reference's 63-bit cell key (what per-cell ascending-id lists imply)."""
    return np.argsort(keys, kind="stable").astype(np.uint32)


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = np.abs(b).max()
    return float(np.abs(a - b).max() / scale) if scale > 0 else float(np.abs(a - b).max())


def cfl_case_inputs(po, scenes, name: str, case: dict):
    """Initial state (params, pos, vel, mass, capacity) of a CFL-limited adaptive-dt case of tests/golden/scalars.json
    (tests/golden/make_golden.py, section 5)."""
    if name == "perf_test":          # the reference benchmark's set-up: initialize(params) + initialize_dam_break(), truncated
        prm = {k: np.float32(v) for k, v in case["params"].items()}
        e = po.Engine("port", case["capacity"]); e.initialize(prm); e.initialize_dam_break()
        s0 = e.state(); e.close()
        assert s0["pos"].shape[0] == case["n"]
        return prm, s0["pos"], None, s0["mass"], case["capacity"]
    pos, mass, prm, _ = scenes.dam_break_scene(0.02)
    vel = None
    if name == "fast_cloud":
        vel = (np.random.default_rng(case["vel_seed"]).normal(size=pos.shape) * case["vel_sigma"]).astype(np.float32)
    else:
        mass = mass.copy()
        mass[0] *= np.float32(case["mass0_factor"])
    return prm, pos, vel, mass, pos.shape[0]
