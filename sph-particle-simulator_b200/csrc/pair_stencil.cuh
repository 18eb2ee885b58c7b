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

__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }
// clamp(fma(a, b, c), 0, 1) in one FFMA.SAT
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// Keeps a kernel parameter in a register for good: ptxas otherwise re-reads it from the constant bank inside the pair
// loops (one issue slot per use).  `zero` is a run-time 0 the compiler cannot see through.
__device__ __forceinline__ float hold(float v, uint32_t zero) { return __uint_as_float(__float_as_uint(v) ^ zero); }

// The density pass of ONE particle over runs of candidates, shared by the per-lane (pair_mask.cu) and the staged
// (pair_stage.cu) kernels so that both produce the same bits.
//   * Candidates are taken two at a time, packed across them on the f32x2 pipe; an odd one at the end of a run takes the
//     same operations in scalar form.
//   * Squared distances carry the reference's roundings (spatial_hash.h:70-73): fl(fl(fl(dx dx) + fl(dy dy)) +
//     fl(dz dz)); the squares are formed as fma(d, d, -0) with a RUN-TIME -0 (an exact product; ptxas would contract a
//     packed multiply with the following packed add into one FFMA2, see pair.cu).
//   * accepted <=> d2 <= r2 <=> d2 - nextafter(r2) < 0: the sign bit, shifted into the column mask with one funnel
//     shift — no compare, no predicate; NaN gives 0 like the reference's compare.
//   * The weight is evaluated unconditionally in compact-support form: W * 6 / (4 sigma) = 2 u^3 - t^3 with
//     u = (2 - q)+ / 2 and t = (1 - q)+, both clamped by the saturating FFMA they are formed with (they lie in [0, 1]):
//     exactly 0 for q >= 2, so there is no "accepted" branch.  TRUNC: the search radius cuts the kernel support short
//     (neighbor_search_radius < 2 h, e.g. the reference's dam-break example: h = 0.025, radius 0.04): the weight is
//     additionally gated by the accept bit.
template <bool TRUNC>
struct DensityWalker {
    float2 npxy, nz2, nr2;
    float npz, nzf, nr2f, ninvh, nhinvh;
    float rho0, rho1;        // sum of m_j W_j * 6 / (4 sigma), even / odd candidates of the runs (the self pair included)
    uint32_t m;              // mask of the column being walked
    unsigned ovf, extra;     // a column held more than 16 candidates; neighbours found beyond the masks

    __device__ __forceinline__ void init(const float4& pi, const PairConsts& k, uint32_t zero) {
        npxy = mk2(-pi.x, -pi.y); npz = -pi.z;
        nzf = hold(k.neg_zero, zero); nz2 = mk2(nzf, nzf);
        nr2f = hold(-k.r2_next, zero); nr2 = mk2(nr2f, nr2f);
        ninvh = hold(-k.inv_h, zero); nhinvh = hold(-0.5f * k.inv_h, zero);
        rho0 = rho1 = 0.0f; m = 0; ovf = 0; extra = 0;
    }
    __device__ __forceinline__ void visit2(const float4& pa, const float4& pb) {
        const float2 da = __fadd2_rn(mk2(pa.x, pa.y), npxy), db = __fadd2_rn(mk2(pb.x, pb.y), npxy);
        const float2 dz = mk2(__fadd_rn(pa.z, npz), __fadd_rn(pb.z, npz));
        const float2 sa = __ffma2_rn(da, da, nz2), sb = __ffma2_rn(db, db, nz2), sz = __ffma2_rn(dz, dz, nz2);
        const float2 d2 = __fadd2_rn(mk2(__fadd_rn(sa.x, sa.y), __fadd_rn(sb.x, sb.y)), sz);
        const float2 t = __fadd2_rn(d2, nr2);
        m = __funnelshift_l(__float_as_uint(t.x), m, 1);
        m = __funnelshift_l(__float_as_uint(t.y), m, 1);
        const float s0 = fast_sqrt(d2.x), s1 = fast_sqrt(d2.y);
        const float2 u = mk2(fma_sat(s0, nhinvh, 1.0f), fma_sat(s1, nhinvh, 1.0f));
        const float2 t1 = mk2(fma_sat(s0, ninvh, 1.0f), fma_sat(s1, ninvh, 1.0f));
        const float2 u3 = __fmul2_rn(__fmul2_rn(u, u), u);
        const float2 nt2 = __fmul2_rn(t1, mk2(-t1.x, -t1.y));
        float2 w = __ffma2_rn(u3, mk2(2.0f, 2.0f), __fmul2_rn(nt2, t1));
        if (TRUNC) {
            w.x = t.x < 0.0f ? w.x : 0.0f;
            w.y = t.y < 0.0f ? w.y : 0.0f;
        }
        rho0 = fmaf(pa.w, w.x, rho0);
        rho1 = fmaf(pb.w, w.y, rho1);
    }
    __device__ __forceinline__ void visit1(const float4& pa) {
        const float2 da = __fadd2_rn(mk2(pa.x, pa.y), npxy);
        const float dz = __fadd_rn(pa.z, npz);
        const float2 sa = __ffma2_rn(da, da, nz2);
        const float d2 = __fadd_rn(__fadd_rn(sa.x, sa.y), __fmaf_rn(dz, dz, nzf));
        const float t = __fadd_rn(d2, nr2f);
        m = __funnelshift_l(__float_as_uint(t), m, 1);
        const float s0 = fast_sqrt(d2);
        const float u = fma_sat(s0, nhinvh, 1.0f), t1 = fma_sat(s0, ninvh, 1.0f);
        float w = __fmaf_rn(__fmul_rn(__fmul_rn(u, u), u), 2.0f, __fmul_rn(__fmul_rn(t1, -t1), t1));
        if (TRUNC) w = t < 0.0f ? w : 0.0f;
        rho0 = fmaf(pa.w, w, rho0);
    }
    // One cell column: candidates [b, e) of `src` (posm itself or a staged copy of a slot range); `to_slot` turns an index
    // of src into a slot of posm.  Returns the 16-bit mask, candidate q at bit 15 - q.  Candidates beyond the mask
    // (collapsed states, coincident wall layers) are walked with the scalar tested loop and set `ovf`; the force pass
    // does the same, so results never depend on the mask capacity.
    __device__ __forceinline__ uint32_t column(const float4* __restrict__ src, uint32_t b, uint32_t e, uint32_t to_slot,
                                               const float4* __restrict__ posm, const float4& pi, const PairConsts& k) {
        const uint32_t end = min(e, b + (uint32_t)kMaskBits);
        m = 0;
        uint32_t j = b;
#pragma unroll 1
        for (; j + 1u < end; j += 2) visit2(src[j], src[j + 1]);
        if (j < end) visit1(src[j]);
        m <<= (b + (uint32_t)kMaskBits) - end;   // end - b candidates were shifted in
        if (e > end) {
            ovf = 1u;
            const float r2 = k.r2, inv_h = k.inv_h;
            for (uint32_t u = end + to_slot; u < e + to_slot; ++u) {
                const float4 pj = __ldg(posm + u);
                const float d2 = dist2_exact(__fsub_rn(pi.x, pj.x), __fsub_rn(pi.y, pj.y), __fsub_rn(pi.z, pj.z));
                if (d2 <= r2) {
                    ++extra;
                    const float q = fast_sqrt(d2) * inv_h;
                    const float uu = fmaxf(1.0f - 0.5f * q, 0.0f), tt = fmaxf(1.0f - q, 0.0f);
                    rho0 += pj.w * (2.0f * (uu * uu * uu) - tt * tt * tt);
                }
            }
        }
        return m;
    }
    __device__ __forceinline__ float density(const PairConsts& k) const { return (rho0 + rho1) * (k.sigma * (4.0f / 6.0f)); }
};

// The force pass of ONE particle over accepted neighbours (no distance test: j was accepted by the density pass), shared
// by the per-lane and the staged kernels so that both produce the same bits.  force_pair_fast (pair_math.cuh) with the
// per-pair constant factors taken out of the sums, r = p_j - p_i (the sign is applied once at the end) and the two clamped
// factors of the spline formed by saturating FFMAs: u = (2 - q)+ / 2, t = (1 - q)+ = sat(2 u - 1) (both in [0, 1]):
//   dW/dq / (2 sigma) = t^2 - u^2,   d2W/dq2 / (2 sigma) = u - 2 t.
// 1/len uses d2 + 1e-30: coincident particles (d2 = 0) get q = 0 and dW/dq(0) = 0, hence no pressure term, exactly like
// the reference's r_len < 1e-6 guard (sph_engine.cpp:403) — the viscosity term keeps L(0) (quirk Q4); a distinct pair
// closer than 1e-6 contributes |dW/dq| <= 2e-6 / h instead of nothing, far below the fast-mode gates.
struct ForceLane {
    float2 npxy, nvxy;
    float pz, vz, P_i, nhinvh;
    // accumulators of -F_pressure / (2 sigma / h) and F_viscosity / (4 mu sigma / h^2): (x, y) packed, z scalar
    float2 fpxy, fvxy;
    float fpz, fvz;

    __device__ __forceinline__ void init(const float4& pi, const float4& vi, float P, const PairConsts& k, uint32_t zero) {
        npxy = mk2(-pi.x, -pi.y); nvxy = mk2(-vi.x, -vi.y);
        pz = pi.z; vz = vi.z; P_i = P;
        nhinvh = hold(-0.5f * k.inv_h, zero);
        fpxy = fvxy = mk2(0.0f, 0.0f);
        fpz = fvz = 0.0f;
    }
    // qa = {x, y, z, A = m / (2 rho)}, qb = {vx, vy, vz, B = A P} of neighbour j
    __device__ __forceinline__ void eval(const float4& qa, const float4& qb) {
        const float2 rxy = __fadd2_rn(mk2(qa.x, qa.y), npxy);
        const float rz = qa.z - pz;
        const float d2 = fmaf(rz, rz, fmaf(rxy.y, rxy.y, rxy.x * rxy.x));
        const float inv_len = fast_rsqrt(d2 + 1e-30f);
        const float u = fma_sat(d2 * nhinvh, inv_len, 1.0f);      // (2 - q)+ / 2
        const float t = fma_sat(u, 2.0f, -1.0f);                   // (1 - q)+
        const float gh = fmaf(t, t, -(u * u));
        const float lq = fmaf(t, -2.0f, u);
        const float cp = fmaf(qa.w, P_i, qb.w) * (gh * inv_len);
        fpxy = __ffma2_rn(mk2(cp, cp), rxy, fpxy);
        fpz = fmaf(cp, rz, fpz);
        const float cv = qa.w * lq;
        const float2 uxy = __fadd2_rn(mk2(qb.x, qb.y), nvxy);
        fvxy = __ffma2_rn(mk2(cv, cv), uxy, fvxy);
        fvz = fmaf(cv, qb.z - vz, fvz);
    }
    // F_p = -sum m_j term gradW with gradW along p_i - p_j = -r and gh = dW/dq / (2 sigma): the two signs cancel
    __device__ __forceinline__ ForceAccum result(const PairConsts& k) const {
        ForceAccum f;
        const float sp = 2.0f * k.sig_h;
        const float cvis = 4.0f * k.viscosity * k.sig_h2;
        f.px = sp * fpxy.x; f.py = sp * fpxy.y; f.pz = sp * fpz;
        f.vx = cvis * fvxy.x; f.vy = cvis * fvxy.y; f.vz = cvis * fvz;
        return f;
    }
};

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
