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
// `edge`: bit 2a / 2a + 1 = the particle sits within a few ulps of the lower / upper face of its cell on axis a: the
// walk then reaches one cell further on that side (see face_bits).
template <typename Body>
__device__ __forceinline__ void walk_runs(const GridDesc& g, const uint32_t* __restrict__ cell_start, int R, int cx, int cy,
                                          int cz, unsigned edge, Body&& body) {
    const int za = max(cz - R - (int)(edge >> 4 & 1u), g.lo[2]), zb = min(cz + R + (int)(edge >> 5 & 1u), g.hi[2]);
    if (za > zb) return;
    const bool split = (za < 0) && (zb >= 0);
    for (int dx = -R - (int)(edge & 1u); dx <= R + (int)(edge >> 1 & 1u); ++dx) {
        const int x = cx + dx;
        if (x < g.lo[0] || x > g.hi[0]) continue;
        const uint32_t bx = (uint32_t)grid_rank(g, 0, x) * (uint32_t)g.ext[1];
        for (int dy = -R - (int)(edge >> 2 & 1u); dy <= R + (int)(edge >> 3 & 1u); ++dy) {
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

// With cells of exactly neighbor_search_radius, two particles that distance apart can sit TWO cells apart when both lie
// within rounding of a cell face (fp32: nsr = 0.03, a = 0.029999996 in cell 0, b = 0.059999995 in cell 2, d2 <= r2);
// the reference finds such a pair because it queries two rings (spatial_hash.cpp:35).  Instead of always walking 125
// cells, a particle within 2^-20 (relative, >= 8 ulps of the product) of a face reaches one cell further on that side
// — the only place such a partner can be.  (Refined grids guard the same tie with 0.1 % larger cells, make_grid.)
__device__ __forceinline__ unsigned face_bits(const float4& p, float inv_cell) {
    unsigned bits = 0;
    const float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float t = __fmul_rn(c[a], inv_cell), f = t - floorf(t), eps = fmaxf(fabsf(t), 1.0f) * 9.5367431640625e-07f;
        if (f <= eps) bits |= 1u << (2 * a);
        if (f >= 1.0f - eps) bits |= 2u << (2 * a);
    }
    return bits;
}

// ---- one thread per particle, private tested walk (strict mode; fast-mode variant 0) -------------------------------------------------
// KT: the smoothing kernel (kKernelCubic = what SPHEngine constructs; Wendland C2 / Gaussian = the other classes behind
// create_kernel, reference kernels.cpp:166-236 — pair_math.cuh)
template <bool STRICT, int KT>
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
        walk_runs(a.grid, a.cell_start, a.walk_radius, cx, cy, cz, face_bits(pi, a.grid.inv_cell), [&](uint32_t b, uint32_t e) {
            SPHB_UNROLL(SPHB_DENSITY_UNROLL)
            for (uint32_t j = b; j < e; ++j) {
                const float4 pj = __ldg(&a.posm[j]);
                const float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y), dz = __fsub_rn(pi.z, pj.z);
                const float d2 = dist2_exact(dx, dy, dz);
                if (d2 <= r2) {
                    ++count;
                    if (STRICT) {
                        if (j != (uint32_t)i) rho = __fadd_rn(rho, __fmul_rn(pj.w, w_strict_k<KT>(a.k, d2)));
                    } else if (KT != kKernelCubic) {
                        rho += pj.w * w_fast_k<KT>(a.k, d2);
                    } else {
                        const float q = fast_sqrt(d2) * inv_h;
                        const float t2 = fmaxf(2.0f - q, 0.0f), t1 = fmaxf(1.0f - q, 0.0f);
                        rho += pj.w * (t2 * t2 * t2 - 4.0f * (t1 * t1 * t1));
                    }
                }
            }
        });
        if (!STRICT) rho *= (KT == kKernelCubic ? sig6 : w_norm_k<KT>(a.k));
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

template <bool STRICT, int KT>
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
    walk_runs(a.grid, a.cell_start, a.walk_radius, cx, cy, cz, face_bits(pi, a.grid.inv_cell), [&](uint32_t b, uint32_t e) {
        SPHB_UNROLL(SPHB_FORCE_UNROLL)
        for (uint32_t j = b; j < e; ++j) {
            const float4 pj = __ldg(&ja[j]);
            const float rx = __fsub_rn(pi.x, pj.x), ry = __fsub_rn(pi.y, pj.y), rz = __fsub_rn(pi.z, pj.z);
            const float d2 = dist2_exact(rx, ry, rz);
            if (d2 <= r2 && (!STRICT || j != (uint32_t)i)) {   // fast: the self pair contributes 0 (r = 0, v_j - v_i = 0)
                const float4 vj = __ldg(&a.fb[j]);
                if (STRICT) {
                    force_pair_strict_k<KT>(a.k, f, rx, ry, rz, d2, __fsub_rn(vj.x, vi.x), __fsub_rn(vj.y, vi.y),
                                            __fsub_rn(vj.z, vi.z), P_i, pj.w, vj.w);
                } else if (KT != kKernelCubic) {
                    force_pair_fast_k<KT>(a.k, f, rx, ry, rz, d2, vj.x - vi.x, vj.y - vi.y, vj.z - vi.z, P_i, pj.w, vj.w);
                } else {
                    force_pair_fast(a.k, f, rx, ry, rz, d2, vj.x - vi.x, vj.y - vi.y, vj.z - vi.z, P_i, pj.w, vj.w);
                }
            }
        }
    });
    a.acc[i] = STRICT ? accel_strict(a.k, f, pi.w) : accel_fast(a.k, f, pi.w);
}

}  // namespace

int launch_density(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    if (a.variant == 2 && !a.strict) return launch_density_mask(a, st);
    const unsigned nb = (unsigned)((a.n + kThreads - 1) / kThreads);
    // strict always takes the scalar reference-order kernel
#define SPHB_LAUNCH_DS(KT) if (a.strict) k_density_simple<true, KT><<<nb, kThreads, 0, st>>>(a); else k_density_simple<false, KT><<<nb, kThreads, 0, st>>>(a)
    switch (a.kernel_type) {
        case kKernelWendlandC2: SPHB_LAUNCH_DS(kKernelWendlandC2); break;
        case kKernelGaussian: SPHB_LAUNCH_DS(kKernelGaussian); break;
        default: SPHB_LAUNCH_DS(kKernelCubic); break;
    }
#undef SPHB_LAUNCH_DS
    return 1;
}

int launch_force(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    if (a.variant == 2 && !a.strict) return launch_force_mask(a, st);
    const unsigned nb = (unsigned)((a.n + kThreads - 1) / kThreads);
#define SPHB_LAUNCH_FS(KT) if (a.strict) k_force_simple<true, KT><<<nb, kThreads, 0, st>>>(a); else k_force_simple<false, KT><<<nb, kThreads, 0, st>>>(a)
    switch (a.kernel_type) {
        case kKernelWendlandC2: SPHB_LAUNCH_FS(kKernelWendlandC2); break;
        case kKernelGaussian: SPHB_LAUNCH_FS(kKernelGaussian); break;
        default: SPHB_LAUNCH_FS(kKernelCubic); break;
    }
#undef SPHB_LAUNCH_FS
    return 1;
}

}  // namespace sphb
