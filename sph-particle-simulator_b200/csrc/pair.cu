// Density (+ equation of state) and force passes over the cell-sorted SoA.
//
// Replaces SPHEngine::update_neighbor_lists' query loop, compute_densities, compute_pressures and
// compute_forces (reference src/sph_engine.cpp:335-353, 203-244): no neighbour lists are ever
// materialised (the reference writes ~1.6 kB of list per particle per step); both passes walk the
// (2R+1)^3 neighbour cells directly in the reference's visiting order — dx, dy, dz ascending, ascending
// particle id inside a cell (src/spatial_hash.cpp:38-52) — so each particle's sums are accumulated in
// the same order as the reference accumulates them over its list.
//
// Because the sort key is (x, y, z)-lexicographic, the 2R+1 cells of one (dx, dy) column are
// contiguous in memory: a particle walks (2R+1)^2 contiguous runs of float4 records (two runs where
// the column crosses the sign change of the masked key, see GridDesc).
#include "pair_math.cuh"

namespace sphb {

namespace {

#ifndef SPHB_PAIR_THREADS
#define SPHB_PAIR_THREADS 128
#endif
#ifndef SPHB_FORCE_UNROLL
#define SPHB_FORCE_UNROLL 4
#endif
#ifndef SPHB_DENSITY_UNROLL
#define SPHB_DENSITY_UNROLL 4
#endif
#define SPHB_STR2(x) #x
#define SPHB_STR(x) SPHB_STR2(x)
#define SPHB_UNROLL(n) _Pragma(SPHB_STR(unroll n))
constexpr int kThreads = SPHB_PAIR_THREADS;
#ifndef SPHB_FORCE_MINBLOCKS
#define SPHB_FORCE_MINBLOCKS 1
#endif
#ifndef SPHB_DENSITY_MINBLOCKS
#define SPHB_DENSITY_MINBLOCKS 1
#endif

// Calls body(begin, end) for every contiguous slot run of the neighbourhood of cell (cx, cy, cz),
// in the reference's visiting order.
template <typename Body>
__device__ __forceinline__ void walk_runs(const GridDesc& g, const uint32_t* __restrict__ cell_start, int R, int cx, int cy,
                                          int cz, Body&& body) {
    const int za = max(cz - R, g.lo[2]), zb = min(cz + R, g.hi[2]);
    if (za > zb) return;
    const bool split = (za < 0) && (zb >= 0);
    for (int dx = -R; dx <= R; ++dx) {
        const int x = cx + dx;
        if (x < g.lo[0] || x > g.hi[0]) continue;
        const uint32_t bx = (uint32_t)grid_rank(g, 0, x) * (uint32_t)g.ext[1];
        for (int dy = -R; dy <= R; ++dy) {
            const int y = cy + dy;
            if (y < g.lo[1] || y > g.hi[1]) continue;
            const uint32_t base = (bx + (uint32_t)grid_rank(g, 1, y)) * (uint32_t)g.ext[2];
            if (!split) {
                body(cell_start[base + grid_rank(g, 2, za)], cell_start[base + grid_rank(g, 2, zb) + 1]);
            } else {
                body(cell_start[base + grid_rank(g, 2, za)], cell_start[base + grid_rank(g, 2, -1) + 1]);
                body(cell_start[base + grid_rank(g, 2, 0)], cell_start[base + grid_rank(g, 2, zb) + 1]);
            }
        }
    }
}

// ---- variant 0: one thread per particle, private walk -------------------------------------------------
template <bool STRICT>
__global__ void __launch_bounds__(kThreads, SPHB_DENSITY_MINBLOCKS) k_density_simple(PairArgs a) {
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    unsigned count = 0;
    if (i < a.n) {
        const float4 pi = a.posm[i];
        const int cx = clampi(cell_coord(pi.x, a.grid.inv_cell), a.grid.lo[0], a.grid.hi[0]);
        const int cy = clampi(cell_coord(pi.y, a.grid.inv_cell), a.grid.lo[1], a.grid.hi[1]);
        const int cz = clampi(cell_coord(pi.z, a.grid.inv_cell), a.grid.lo[2], a.grid.hi[2]);
        // density = m_i * W(0)   (sph_engine.cpp:373)
        // strict: density = m_i * W(0) first, then the j != i terms in list order (sph_engine.cpp:373-380).
        // fast: the self pair (d2 = 0) stays in the loop — the polynomial gives exactly sigma * 4/6 there —
        // which removes the j != i test from the hot loop.
        float rho = STRICT ? __fmul_rn(pi.w, a.k.w0) : 0.0f;
        const float r2 = a.k.r2;
        const float inv_h = a.k.inv_h, sig6 = a.k.sigma * (1.0f / 6.0f);
        walk_runs(a.grid, a.cell_start, a.walk_radius, cx, cy, cz, [&](uint32_t b, uint32_t e) {
            SPHB_UNROLL(SPHB_DENSITY_UNROLL)
            for (uint32_t j = b; j < e; ++j) {
                const float4 pj = __ldg(&a.posm[j]);
                const float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y), dz = __fsub_rn(pi.z, pj.z);
                const float d2 = dist2_exact(dx, dy, dz);
                if (d2 <= r2) {
                    ++count;
                    if (STRICT) {
                        if (j != (uint32_t)i) rho = __fadd_rn(rho, __fmul_rn(pj.w, w_strict(a.k, d2)));
                    } else {
                        const float q = fast_sqrt(d2) * inv_h;
                        const float t2 = fmaxf(2.0f - q, 0.0f), t1 = fmaxf(1.0f - q, 0.0f);
                        rho += pj.w * (t2 * t2 * t2 - 4.0f * (t1 * t1 * t1));
                    }
                }
            }
        });
        if (!STRICT) rho *= sig6;
        float P;
        if (STRICT) P = __fmul_rn(a.k.gas_constant, __fsub_rn(rho, a.k.rest_density));
        else P = a.k.gas_constant * (rho - a.k.rest_density);
        a.rho_p[i] = make_float2(rho, P);
        const float4 v = a.velid[i];
        if (STRICT) {
            a.fb[i] = make_float4(v.x, v.y, v.z, rho);
        } else {
            const float A = pi.w / (2.0f * rho);
            a.fa[i] = make_float4(pi.x, pi.y, pi.z, A);
            a.fb[i] = make_float4(v.x, v.y, v.z, A * P);
        }
        if (a.nbr_count) a.nbr_count[i] = count;
    }
    count = __reduce_max_sync(0xffffffffu, count);
    if ((threadIdx.x & 31) == 0 && count > *(volatile unsigned int*)&a.sc->max_neighbors) atomicMax(&a.sc->max_neighbors, count);
}

template <bool STRICT>
__global__ void __launch_bounds__(kThreads, SPHB_FORCE_MINBLOCKS) k_force_simple(PairArgs a) {
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= a.n) return;
    const float4 pi = a.posm[i];
    const float4 vi = a.velid[i];
    if (is_ghost(vi)) return;   // slab mode: halo copies are never advanced here
    const float P_i = a.rho_p[i].y;
    const int cx = clampi(cell_coord(pi.x, a.grid.inv_cell), a.grid.lo[0], a.grid.hi[0]);
    const int cy = clampi(cell_coord(pi.y, a.grid.inv_cell), a.grid.lo[1], a.grid.hi[1]);
    const int cz = clampi(cell_coord(pi.z, a.grid.inv_cell), a.grid.lo[2], a.grid.hi[2]);
    ForceAccum f = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const float r2 = a.k.r2;
    const float4* __restrict__ ja = STRICT ? a.posm : a.fa;
    walk_runs(a.grid, a.cell_start, a.walk_radius, cx, cy, cz, [&](uint32_t b, uint32_t e) {
        SPHB_UNROLL(SPHB_FORCE_UNROLL)
        for (uint32_t j = b; j < e; ++j) {
            const float4 pj = __ldg(&ja[j]);
            const float rx = __fsub_rn(pi.x, pj.x), ry = __fsub_rn(pi.y, pj.y), rz = __fsub_rn(pi.z, pj.z);
            const float d2 = dist2_exact(rx, ry, rz);
            if (d2 <= r2 && (!STRICT || j != (uint32_t)i)) {   // fast: the self pair contributes 0 (r = 0, v_j - v_i = 0)
                const float4 vj = __ldg(&a.fb[j]);
                if (STRICT) {
                    force_pair_strict(a.k, f, rx, ry, rz, d2, __fsub_rn(vj.x, vi.x), __fsub_rn(vj.y, vi.y),
                                      __fsub_rn(vj.z, vi.z), P_i, pj.w, vj.w);
                } else {
                    force_pair_fast(a.k, f, rx, ry, rz, d2, vj.x - vi.x, vj.y - vi.y, vj.z - vi.z, P_i, pj.w, vj.w);
                }
            }
        }
    });
    a.acc[i] = STRICT ? accel_strict(a.k, f, pi.w) : accel_fast(a.k, f, pi.w);
}

// ---- variant 1 (fast math only): packed f32x2 arithmetic, two candidates per instruction -----------------
//
// ncu on variant 0 (profiles/): both pair kernels are ISSUE-bound (≈85-90 % issue-slot utilisation,
// <1 % DRAM), ≈28 / 41 warp-instructions per candidate.  sm_100 adds packed fp32 instructions
// (FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per lane per issue slot), so the lever is to put TWO
// candidates into every arithmetic instruction:
//   * the sorted state is mirrored into a pair-interleaved layout {x0,x1,y0,y1 | z0,z1,w0,w1} (32 B per
//     two particles, written by the reorder kernel), so two LDG.128 deliver register pairs that are
//     already aligned for the packed ALU ops — no shuffles or moves;
//   * the pair arithmetic is evaluated unconditionally in branch-free B-spline form and masked through
//     q (q := 2 ⇒ every kernel term is exactly 0), which removes the divergent "accepted" branch;
//   * the radius test stays the exact one: squares are formed as fma(d, d, -0) (= the correctly rounded
//     product; ptxas would otherwise fuse mul.rn.f32x2 + add.rn.f32x2 into FFMA2) and summed with two
//     separately rounded adds, so neighbour sets and counts are bit-identical to the reference's.
// Even and odd candidates accumulate in the two halves of packed accumulators (summation order differs
// from the reference ⇒ fast mode only; strict mode always runs variant 0).
typedef float2 f2;

__device__ __forceinline__ f2 f2_set(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// Correctly rounded a*a as fma(a, a, -0.0).  The addend must be opaque to the compiler: with a literal
// -0.0 NVVM folds the fma back into a multiply, and ptxas (12.9) then contracts mul.rn.f32x2 +
// add.rn.f32x2 into FFMA2 even though both carry .rn — which would make the radius test inexact.
// nz holds -0.0f loaded from the kernel parameters.
__device__ __forceinline__ f2 f2_sq_exact(f2 a, f2 nz) { return __ffma2_rn(a, a, nz); }
__device__ __forceinline__ f2 f2_max0(f2 a) { return make_float2(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f)); }

struct PairRec {   // two consecutive sorted particles
    f2 x, y, z, w;
};

__device__ __forceinline__ PairRec load_pair(const float4* __restrict__ base, uint32_t k) {
    const float4 a = __ldg(&base[2 * (size_t)k]);
    const float4 b = __ldg(&base[2 * (size_t)k + 1]);
    PairRec r;
    r.x = make_float2(a.x, a.y); r.y = make_float2(a.z, a.w);
    r.z = make_float2(b.x, b.y); r.w = make_float2(b.z, b.w);
    return r;
}

// store one particle's 4 values into its half of pair record s >> 1
__device__ __forceinline__ void store_half(float4* base, size_t s, float x, float y, float z, float w) {
    float* f = reinterpret_cast<float*>(base) + 8 * (s >> 1) + (s & 1);
    f[0] = x; f[2] = y; f[4] = z; f[6] = w;
}

template <int R_UNUSED = 0>
__global__ void __launch_bounds__(kThreads) k_density_packed(PairArgs a) {
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    unsigned count = 0;
    if (i < a.n && wants_density(a, a.posm[i])) {
        const float4 pi = a.posm[i];
        const int cx = clampi(cell_coord(pi.x, a.grid.inv_cell), a.grid.lo[0], a.grid.hi[0]);
        const int cy = clampi(cell_coord(pi.y, a.grid.inv_cell), a.grid.lo[1], a.grid.hi[1]);
        const int cz = clampi(cell_coord(pi.z, a.grid.inv_cell), a.grid.lo[2], a.grid.hi[2]);
        const f2 px = f2_set(pi.x), py = f2_set(pi.y), pz = f2_set(pi.z);
        const f2 inv_h = f2_set(a.k.inv_h);
        const f2 nz = f2_set(a.k.neg_zero);
        const float r2 = a.k.r2;
        f2 acc = f2_set(0.0f);
        const uint32_t self = (uint32_t)i;
        walk_runs(a.grid, a.cell_start, a.walk_radius, cx, cy, cz, [&](uint32_t b, uint32_t e) {
            if (b >= e) return;
            const uint32_t k1 = (e - 1) >> 1;
#pragma unroll 2
            for (uint32_t k = b >> 1; k <= k1; ++k) {
                const PairRec c = load_pair(a.pp2, k);
                const f2 dx = f2_sub(px, c.x), dy = f2_sub(py, c.y), dz = f2_sub(pz, c.z);
                const f2 d2 = __fadd2_rn(__fadd2_rn(f2_sq_exact(dx, nz), f2_sq_exact(dy, nz)), f2_sq_exact(dz, nz));
                const uint32_t j0 = 2 * k, j1 = 2 * k + 1;
                const bool in0 = (d2.x <= r2) && (j0 >= b) && (j0 < e);
                const bool in1 = (d2.y <= r2) && (j1 >= b) && (j1 < e);
                count += (in0 ? 1u : 0u) + (in1 ? 1u : 0u);
                f2 q = __fmul2_rn(make_float2(fast_sqrt(d2.x), fast_sqrt(d2.y)), inv_h);
                q.x = (in0 && j0 != self) ? q.x : 2.0f;
                q.y = (in1 && j1 != self) ? q.y : 2.0f;
                const f2 t2 = f2_max0(f2_sub(f2_set(2.0f), q));
                const f2 t1 = f2_max0(f2_sub(f2_set(1.0f), q));
                const f2 t2c = __fmul2_rn(__fmul2_rn(t2, t2), t2);
                const f2 t1c = __fmul2_rn(__fmul2_rn(t1, t1), t1);
                const f2 poly = __ffma2_rn(f2_set(-4.0f), t1c, t2c);      // (2-q)+^3 - 4 (1-q)+^3
                acc = __ffma2_rn(poly, c.w, acc);
            }
        });
        // rho = m_i W(0) + sigma/6 * sum m_j [(2-q)+^3 - 4 (1-q)+^3]
        const float rho = pi.w * a.k.w0 + (a.k.sigma * (1.0f / 6.0f)) * (acc.x + acc.y);
        const float P = a.k.gas_constant * (rho - a.k.rest_density);
        a.rho_p[i] = make_float2(rho, P);
        const float4 v = a.velid[i];
        // force-pass inputs, folded once per particle: A' = (sigma/h) m / (2 rho), B' = A' P
        const float A = a.k.sig_h * (pi.w / (2.0f * rho));
        store_half(a.fa2, i, pi.x, pi.y, pi.z, A);
        store_half(a.fb2, i, v.x, v.y, v.z, A * P);
        if (a.nbr_count) a.nbr_count[i] = count;
    }
    count = __reduce_max_sync(0xffffffffu, count);
    if ((threadIdx.x & 31) == 0 && count > *(volatile unsigned int*)&a.sc->max_neighbors) atomicMax(&a.sc->max_neighbors, count);
}

template <int R_UNUSED = 0>
__global__ void __launch_bounds__(kThreads) k_force_packed(PairArgs a) {
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= a.n) return;
    const float4 pi = a.posm[i];
    const float4 vi = a.velid[i];
    if (is_ghost(vi)) return;
    const int cx = clampi(cell_coord(pi.x, a.grid.inv_cell), a.grid.lo[0], a.grid.hi[0]);
    const int cy = clampi(cell_coord(pi.y, a.grid.inv_cell), a.grid.lo[1], a.grid.hi[1]);
    const int cz = clampi(cell_coord(pi.z, a.grid.inv_cell), a.grid.lo[2], a.grid.hi[2]);
    const f2 px = f2_set(pi.x), py = f2_set(pi.y), pz = f2_set(pi.z);
    const f2 ux = f2_set(vi.x), uy = f2_set(vi.y), uz = f2_set(vi.z);
    const f2 P_i = f2_set(a.rho_p[i].y);
    const f2 inv_h = f2_set(a.k.inv_h);
    const f2 nz = f2_set(a.k.neg_zero);
    const f2 cvis = f2_set(2.0f * a.k.viscosity * a.k.inv_h);   // 2 mu (sigma/h^2) / (sigma/h)
    const float r2 = a.k.r2;
    f2 fpx = f2_set(0.0f), fpy = f2_set(0.0f), fpz = f2_set(0.0f);
    f2 fvx = f2_set(0.0f), fvy = f2_set(0.0f), fvz = f2_set(0.0f);
    const uint32_t self = (uint32_t)i;
    walk_runs(a.grid, a.cell_start, a.walk_radius, cx, cy, cz, [&](uint32_t b, uint32_t e) {
        if (b >= e) return;
        const uint32_t k1 = (e - 1) >> 1;
#pragma unroll 2
        for (uint32_t k = b >> 1; k <= k1; ++k) {
            const PairRec c = load_pair(a.fa2, k);
            const PairRec v = load_pair(a.fb2, k);
            const f2 rx = f2_sub(px, c.x), ry = f2_sub(py, c.y), rz = f2_sub(pz, c.z);
            const f2 d2 = __fadd2_rn(__fadd2_rn(f2_sq_exact(rx, nz), f2_sq_exact(ry, nz)), f2_sq_exact(rz, nz));
            const uint32_t j0 = 2 * k, j1 = 2 * k + 1;
            const bool in0 = (d2.x <= r2) && (j0 >= b) && (j0 < e) && (j0 != self);
            const bool in1 = (d2.y <= r2) && (j1 >= b) && (j1 < e) && (j1 != self);
            // 1/len with d2 clamped away from 0: coincident particles get q = 0 and a pressure term
            // (dW/dq)(0) / len * r = 0 * r = 0, the viscosity term keeps its L(0) (reference Q4)
            const f2 inv_len = make_float2(fast_rsqrt(fmaxf(d2.x, 1e-30f)), fast_rsqrt(fmaxf(d2.y, 1e-30f)));
            f2 q = __fmul2_rn(__fmul2_rn(d2, inv_len), inv_h);
            q.x = in0 ? q.x : 2.0f;
            q.y = in1 ? q.y : 2.0f;
            const f2 t2 = f2_max0(f2_sub(f2_set(2.0f), q));
            const f2 t1 = f2_max0(f2_sub(f2_set(1.0f), q));
            // dW/dq / sigma = 2 (1-q)+^2 - 0.5 (2-q)+^2 ;  d2W/dq2 / sigma = (2-q)+ - 4 (1-q)+
            const f2 gq = __ffma2_rn(f2_set(-0.5f), __fmul2_rn(t2, t2), __fmul2_rn(f2_set(2.0f), __fmul2_rn(t1, t1)));
            const f2 lq = __ffma2_rn(f2_set(-4.0f), t1, t2);
            // pressure: F -= m_j (P_i + P_j) / (2 rho_j) * (sigma/h) gq * r / len
            const f2 cp = __fmul2_rn(__ffma2_rn(c.w, P_i, v.w), __fmul2_rn(gq, inv_len));
            fpx = __ffma2_rn(make_float2(-cp.x, -cp.y), rx, fpx);
            fpy = __ffma2_rn(make_float2(-cp.x, -cp.y), ry, fpy);
            fpz = __ffma2_rn(make_float2(-cp.x, -cp.y), rz, fpz);
            // viscosity: F += (m_j / rho_j) mu (v_j - v_i) (sigma/h^2) lq
            const f2 cv = __fmul2_rn(__fmul2_rn(cvis, c.w), lq);
            fvx = __ffma2_rn(cv, f2_sub(v.x, ux), fvx);
            fvy = __ffma2_rn(cv, f2_sub(v.y, uy), fvy);
            fvz = __ffma2_rn(cv, f2_sub(v.z, uz), fvz);
        }
    });
    ForceAccum f;
    f.px = fpx.x + fpx.y; f.py = fpy.x + fpy.y; f.pz = fpz.x + fpz.y;
    f.vx = fvx.x + fvx.y; f.vy = fvy.x + fvy.y; f.vz = fvz.x + fvz.y;
    a.acc[i] = accel_fast(a.k, f, pi.w);
}

}  // namespace

int launch_density(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    if (a.variant == 2 && !a.strict) return launch_density_mask(a, st);
    const unsigned nb = (unsigned)((a.n + kThreads - 1) / kThreads);
    if (a.variant == 0) {
        if (a.strict) k_density_simple<true><<<nb, kThreads, 0, st>>>(a);
        else k_density_simple<false><<<nb, kThreads, 0, st>>>(a);
    } else {
        if (a.strict) k_density_simple<true><<<nb, kThreads, 0, st>>>(a);   // strict always takes the scalar reference-order kernel
        else k_density_packed<0><<<nb, kThreads, 0, st>>>(a);
    }
    return 1;
}

int launch_force(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    if (a.variant == 2 && !a.strict) return launch_force_mask(a, st);
    const unsigned nb = (unsigned)((a.n + kThreads - 1) / kThreads);
    if (a.variant == 0) {
        if (a.strict) k_force_simple<true><<<nb, kThreads, 0, st>>>(a);
        else k_force_simple<false><<<nb, kThreads, 0, st>>>(a);
    } else {
        if (a.strict) k_force_simple<true><<<nb, kThreads, 0, st>>>(a);
        else k_force_packed<0><<<nb, kThreads, 0, st>>>(a);
    }
    return 1;
}

}  // namespace sphb
