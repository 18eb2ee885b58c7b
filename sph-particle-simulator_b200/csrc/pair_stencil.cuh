// Static spherical stencil, cell addressing and record loads shared by the bitmask pair kernels
// (pair_mask.cu: 16-bit column masks, R >= 4; pair_mask_wide.cu: 64-bit column masks, R = 2, 3).
//
// Reference: the 125-cell query of SpatialHash::query_squared (src/spatial_hash.cpp:31-57); the stencil below is a
// superset of the cells that can hold an accepted neighbour on the refined grid (tests/test_stencil_model.py).
#pragma once

#include "pair_math.cuh"

namespace sphb {

namespace {

// A particle in cell c can only have neighbours (distance <= nsr <= R cells) in cells whose offset (d0, d1, d2)
// satisfies (|d0|-1)+^2 + (|d1|-1)+^2 + (|d2|-1)+^2 <= R^2 (the gap between two cells is at least |d|-1 cell widths),
// so column (d0, d1) needs the cells |d2| <= reach(d0, d1) only — independent of the lane, hence free of divergence.
// R = 4: 613 of the 729 cells (4 corner columns drop out entirely), R = 3: 335 of 343, R = 2: all 125.
// reach = -1: the column cannot contain a neighbour.  Tables in walk order (d0 outer, d1 inner).
struct ReachTables {
    int r2[25], r3[49], r4[81], r5[121], r6[169];   // int, not char: a uniform 32-bit constant load with no sign extension
};
constexpr int isqrt_floor(int v) {
    int r = 0;
    while ((r + 1) * (r + 1) <= v) ++r;
    return r;
}
constexpr int reach_of(int R, int d0, int d1) {
    const int a0 = (d0 < 0 ? -d0 : d0) - 1, a1 = (d1 < 0 ? -d1 : d1) - 1;
    const int rem = R * R - (a0 > 0 ? a0 * a0 : 0) - (a1 > 0 ? a1 * a1 : 0);
    if (rem < 0) return -1;
    const int z = isqrt_floor(rem) + 1;
    return z < R ? z : R;
}
constexpr ReachTables make_reach_tables() {
    ReachTables t{};
    for (int d0 = -2; d0 <= 2; ++d0) for (int d1 = -2; d1 <= 2; ++d1) t.r2[(d0 + 2) * 5 + d1 + 2] = reach_of(2, d0, d1);
    for (int d0 = -3; d0 <= 3; ++d0) for (int d1 = -3; d1 <= 3; ++d1) t.r3[(d0 + 3) * 7 + d1 + 3] = reach_of(3, d0, d1);
    for (int d0 = -4; d0 <= 4; ++d0) for (int d1 = -4; d1 <= 4; ++d1) t.r4[(d0 + 4) * 9 + d1 + 4] = reach_of(4, d0, d1);
    for (int d0 = -5; d0 <= 5; ++d0) for (int d1 = -5; d1 <= 5; ++d1) t.r5[(d0 + 5) * 11 + d1 + 5] = reach_of(5, d0, d1);
    for (int d0 = -6; d0 <= 6; ++d0) for (int d1 = -6; d1 <= 6; ++d1) t.r6[(d0 + 6) * 13 + d1 + 6] = reach_of(6, d0, d1);
    return t;
}
__constant__ ReachTables kReach = make_reach_tables();

template <int R>
__device__ __forceinline__ int column_reach(int col) {
    return R == 2 ? kReach.r2[col] : (R == 3 ? kReach.r3[col] : (R == 4 ? kReach.r4[col] : (R == 5 ? kReach.r5[col] : kReach.r6[col])));
}

// Linear index of the particle's cell in the (padded) cell table.  The fast-mode grid carries g.pad >= R empty cells
// around the populated box on every axis, and the cell is clamped into the populated box exactly like the key kernel
// does, so every cell of every column of the stencil exists in the table: no range checks in the walks.
__device__ __forceinline__ uint32_t center_cell(const GridDesc& g, const float4& p) {
    const int c0 = clampi(cell_coord(pick_axis(p, g.perm[0]), g.inv_cell), g.lo[0] + g.pad, g.hi[0] - g.pad) - g.lo[0];
    const int c1 = clampi(cell_coord(pick_axis(p, g.perm[1]), g.inv_cell), g.lo[1] + g.pad, g.hi[1] - g.pad) - g.lo[1];
    const int c2 = clampi(cell_coord(pick_axis(p, g.perm[2]), g.inv_cell), g.lo[2] + g.pad, g.hi[2] - g.pad) - g.lo[2];
    return ((uint32_t)c0 * (uint32_t)g.ext[1] + (uint32_t)c1) * (uint32_t)g.ext[2] + (uint32_t)c2;
}

// ---- 16-bit column masks (pair_mask.cu, pair_stage.cu; R >= 4) ----
constexpr int kMaskBits = 16;   // candidates per column mask

// Mask rows: one per mirror-pair group whose columns lie inside the stencil, in walk order, plus the centre column.
// Bit 31 of the centre word is the particle's overflow flag (the centre column only uses the low half).
template <int R>
struct Groups {
    static constexpr int kCols = (2 * R + 1) * (2 * R + 1);
    static constexpr int kGroups = kCols / 2;   // groups 0 .. kGroups-1 are {column g, column kCols-1-g}, group kGroups is the centre
};

// The non-empty mirror-pair groups of the stencil in walk order, then the centre column (d0 = d1 = 0) as entry n.
template <int R>
struct GroupTable {
    int n;
    int d0[Groups<R>::kGroups + 1], d1[Groups<R>::kGroups + 1], reach[Groups<R>::kGroups + 1];
};
template <int R>
constexpr GroupTable<R> make_group_table() {
    GroupTable<R> t{};
    const int w = 2 * R + 1;
    for (int g = 0; g < Groups<R>::kGroups; ++g) {
        const int r = reach_of(R, g / w - R, g % w - R);
        if (r >= 0) { t.d0[t.n] = g / w - R; t.d1[t.n] = g % w - R; t.reach[t.n] = r; ++t.n; }
    }
    t.d0[t.n] = 0; t.d1[t.n] = 0; t.reach[t.n] = reach_of(R, 0, 0);
    return t;
}
__constant__ GroupTable<4> kGroups4 = make_group_table<4>();
__constant__ GroupTable<5> kGroups5 = make_group_table<5>();
__constant__ GroupTable<6> kGroups6 = make_group_table<6>();
template <int R> __device__ __forceinline__ const GroupTable<R>& group_table();
template <> __device__ __forceinline__ const GroupTable<4>& group_table<4>() { return kGroups4; }
template <> __device__ __forceinline__ const GroupTable<5>& group_table<5>() { return kGroups5; }
template <> __device__ __forceinline__ const GroupTable<6>& group_table<6>() { return kGroups6; }

// 32-byte force-pass record of one particle, fetched with ONE 256-bit load (LDG.E.256, sm_100)
__device__ __forceinline__ ForceRec load_rec(const ForceRec* __restrict__ p) {
    ForceRec r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.A), "=f"(r.vx), "=f"(r.vy), "=f"(r.vz), "=f"(r.B)
        : "l"(p));
    return r;
}

}  // namespace

}  // namespace sphb
