"""CPU: the fast path's static spherical stencil (csrc/pair_mask.cu, exported by sphb_debug_stencil) never loses a
neighbour.  A numpy model repeats the device's fp32 cell assignment — cell = floor(fl(p * inv_cell)) with
inv_cell = fl(fl(1 / nsr) * refine) * (1 - 2^-10), csrc/api.cu make_grid — and checks, for every pair the reference
accepts ((dx*dx + dy*dy) + dz*dz <= nsr*nsr in fp32, spatial_hash.h:70-73), that the partner's cell lies inside the
stencil of the particle's cell.  Includes lattices whose spacing puts many pairs EXACTLY at distance nsr (the q = 2
ties of the dam-break scenes), which is what the 0.1 % cell margin is for."""
import numpy as np
import pytest

f32 = np.float32


def accepted_pairs(pos, nsr):
    """Brute force in fp32 with the reference's association order; returns index arrays (i, j), i != j."""
    r2 = f32(nsr) * f32(nsr)
    out_i, out_j = [], []
    for i in range(len(pos)):
        d = pos[i][None, :] - pos                      # fp32 subtractions p_i - p_j
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        j = np.flatnonzero(d2 <= r2)
        j = j[j != i]
        out_i.append(np.full(len(j), i)); out_j.append(j)
    return np.concatenate(out_i), np.concatenate(out_j)


def clouds():
    rng = np.random.default_rng(5)
    dx = 0.004
    g = np.arange(14, dtype=np.float32) * f32(dx)
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    yield "lattice_ties", lattice - f32(0.02), 4 * dx                       # nsr = 4 dx: exact ties along the axes
    yield "lattice_negative_offset", lattice - f32(0.0391), 4 * dx
    yield "lattice_h25", lattice + f32(0.3), 5 * dx                           # h = 2.5 dx family
    yield "random", rng.uniform(-0.03, 0.03, size=(1500, 3)).astype(np.float32), 0.016
    yield "random_far_from_origin", (rng.uniform(0, 0.05, size=(1200, 3)) + 3.0).astype(np.float32), 0.016


@pytest.mark.parametrize("radius", [2, 3, 4, 5, 6])
def test_stencil_covers_every_accepted_pair(pkg, radius):
    reach, scale = pkg.capi.debug_stencil(radius)
    assert reach.shape == (2 * radius + 1, 2 * radius + 1)
    assert (reach == reach[::-1, ::-1]).all() and (reach == reach.T).all()      # mirror columns share their reach
    assert reach[radius, radius] == radius and reach.max() == radius
    for name, pos, nsr in clouds():
        inv_cell = f32(f32(f32(1.0) / f32(nsr)) * f32(radius)) * f32(scale)       # walk radius 1 x refine = radius
        cell = np.floor(pos * inv_cell).astype(np.int64)                          # __float2int_rd(__fmul_rn(p, inv_cell))
        i, j = accepted_pairs(pos, nsr)
        assert len(i) > 1000, name
        d = cell[j] - cell[i]
        assert (np.abs(d) <= radius).all(), f"{name}: partner outside the {2 * radius + 1}^3 cube"
        r = reach[d[:, 0] + radius, d[:, 1] + radius]
        assert (r >= 0).all(), f"{name}: accepted pair in a skipped column"
        assert (np.abs(d[:, 2]) <= r).all(), f"{name}: accepted pair beyond the column reach"


def test_stencil_is_minimal_enough():
    """Documented cell counts of the stencil: 125, 335, 613, 1 087, 1 713 cells for R = 2..6."""
    import __graft_entry__ as g
    capi = g.load_package().capi
    want = {2: 125, 3: 335, 4: 613, 5: 1087, 6: 1713}
    for R, cells in want.items():
        reach, _ = capi.debug_stencil(R)
        assert int((2 * reach[reach >= 0].astype(int) + 1).sum()) == cells
