// Pair kernels of the fast path for SMALL particle sets: L lanes share one particle (SPHB_OPT_LANES_PER_PARTICLE).
//
// With one thread per particle a step of a small scene is bound by the serial walk of a single thread — ~560 candidates
// in the density pass, ~240 pops in the force pass — while most of the machine idles: the reference's own drivers run
// 1 000 – 20 000 particles (benchmarks/performance_test.cpp:131-140, examples/dam_break.cpp:40-61), i.e. 13 k threads on
// a 303 k-thread device.  Here the mirror-pair column groups of a particle are dealt round-robin to L adjacent lanes
// (lane s takes groups s, s + L, ...), each lane runs the unchanged DensityWalker / ForceLane code on its share, and the
// partial sums meet in L-lane shuffle reductions.  Same neighbour sets, same masks, same per-pair arithmetic as
// pair_mask.cu; only the association order of the per-particle sums differs (fast-mode tolerances), so the
// variant is selected explicitly and never mixed with the one-lane kernels inside one comparison.
//
// Reference: SPHEngine::update_neighbor_lists' query + compute_densities + compute_pressures + compute_forces
// (src/sph_engine.cpp:335-353, 203-244).
#include "pair_stencil.cuh"

namespace sphb {

namespace {

constexpr int kThreadsP = 128;

__device__ __forceinline__ uint32_t top_bit(uint32_t w) {
    uint32_t b;
    asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(w));
    return b;
}
__device__ __forceinline__ uint32_t bit_at(uint32_t pos) {
    uint32_t m;
    asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(m) : "r"(pos));
    return m;
}

template <bool SLAB, int R, bool TRUNC, int L>
__global__ void __launch_bounds__(kThreadsP) k_density_split(PairArgs a) {
    __shared__ int4 soff[Groups<R>::kGroups + 1];   // per group: cell offsets of {run start, run end} of column and mirror
    const size_t t = (size_t)blockIdx.x * kThreadsP + threadIdx.x;
    const unsigned s = (unsigned)(t % L);            // this lane's share: groups s, s + L, ...
    const size_t i_raw = t / L;
    const size_t i = i_raw < a.n ? i_raw : a.n - 1;
    const float4 pi = a.posm[i];
    const bool want = i_raw < a.n && (!SLAB || wants_density(a, pi));   // uniform over the L lanes of a particle
    const GroupTable<R>& tab = group_table<R>();
    const int ng = tab.n;
    {
        const int e2 = a.grid.ext[2], e12 = a.grid.ext[1] * e2;
        for (int g = threadIdx.x; g <= ng; g += kThreadsP) {
            const int rel = tab.d0[g] * e12 + tab.d1[g] * e2, reach = tab.reach[g];
            soff[g] = make_int4(rel - reach, rel + reach + 1, -rel - reach, -rel + reach + 1);
        }
    }
    __syncthreads();
    const unsigned peers = __ballot_sync(0xffffffffu, want);   // the L lanes of a particle enter together
    unsigned count = 0;
    if (want) {
        const uint32_t* __restrict__ csc = a.cell_start + center_cell(a.grid, pi);
        const float4* __restrict__ posm = a.posm;
        DensityWalker<TRUNC> dw;
        dw.init(pi, a.k, (uint32_t)(a.n >> 62));
        uint32_t* __restrict__ mrow = static_cast<uint32_t*>(a.masks) + i;
        const size_t stride = a.mask_stride;
        for (int g = (int)s; g < ng; g += L) {
            const int4 o = soff[g];
            const uint32_t bA = __ldg(csc + o.x), eA = __ldg(csc + o.y), bB = __ldg(csc + o.z), eB = __ldg(csc + o.w);
            uint32_t word = dw.column(posm, bA, eA, 0u, posm, pi, a.k);
            word |= dw.column(posm, bB, eB, 0u, posm, pi, a.k) << 16;
            count += __popc(word);
            __stcs(mrow + (size_t)g * stride, word);
        }
        uint32_t centre = 0;
        const bool mine = s == (unsigned)(ng % L);   // the centre column goes to the lane whose turn it is
        if (mine) {
            const int4 o = soff[ng];
            centre = dw.column(posm, __ldg(csc + o.x), __ldg(csc + o.y), 0u, posm, pi, a.k);
            count += __popc(centre);
        }
        count += dw.extra;
        float part = dw.rho0 + dw.rho1;
        unsigned ovf = dw.ovf;
#pragma unroll
        for (int off = 1; off < L; off <<= 1) {
            part += __shfl_xor_sync(peers, part, off);
            count += __shfl_xor_sync(peers, count, off);
            ovf |= __shfl_xor_sync(peers, ovf, off);
        }
        if (mine) __stcs(mrow + (size_t)ng * stride, centre | (ovf << 31));
        if (s == 0) {
            const float rho = part * (a.k.sigma * (4.0f / 6.0f));
            const float P = a.k.gas_constant * (rho - a.k.rest_density);
            a.rho_p[i] = make_float2(rho, P);
            const float4 v = a.velid[i];
            const float A = pi.w / (2.0f * rho);
            a.fa[i] = make_float4(pi.x, pi.y, pi.z, A);
            a.fb[i] = make_float4(v.x, v.y, v.z, A * P);
            if (a.nbr_count) a.nbr_count[i] = count;
        }
    }
    count = __reduce_max_sync(0xffffffffu, count);
    if ((threadIdx.x & 31) == 0 && count > *(volatile unsigned int*)&a.sc->max_neighbors) atomicMax(&a.sc->max_neighbors, count);
}

template <bool SLAB, int R, int L>
__global__ void __launch_bounds__(kThreadsP) k_force_split(PairArgs a) {
    const size_t t = (size_t)blockIdx.x * kThreadsP + threadIdx.x;
    const unsigned s = (unsigned)(t % L);
    const size_t i_raw = t / L;
    const size_t i = i_raw < a.n ? i_raw : a.n - 1;
    const float4 vi = a.velid[i];
    const bool want = i_raw < a.n && !(SLAB && is_ghost(vi));   // slab mode: halo copies are never advanced here
    const unsigned peers = __ballot_sync(0xffffffffu, want);
    if (!want) return;
    const float4 pi = a.posm[i];
    const float P_i = a.rho_p[i].y;
    const uint32_t c = center_cell(a.grid, pi);
    const uint32_t* __restrict__ cs = a.cell_start;
    const float4* __restrict__ fa = a.fa;
    const float4* __restrict__ fb = a.fb;
    ForceLane fl;
    fl.init(pi, vi, P_i, a.k, (uint32_t)(a.n >> 62));
    const GroupTable<R>& tab = group_table<R>();
    const int ng = tab.n;
    const int e2 = a.grid.ext[2], e12 = a.grid.ext[1] * e2;
    const uint32_t* __restrict__ mrow = static_cast<const uint32_t*>(a.masks) + i;
    const size_t stride = a.mask_stride;
    unsigned ovf = 0;
    for (int g = (int)s; g <= ng; g += L) {
        uint32_t w = __ldcs(mrow + (size_t)g * stride);
        if (g == ng) { ovf = w >> 31; w &= 0xFFFFu; }
        const uint32_t rel = (uint32_t)(tab.d0[g] * e12 + tab.d1[g] * e2), reach = (uint32_t)tab.reach[g];
        // candidate q of the group's column is bit 15 - q, of its mirror bit 31 - q: slot = base - bit
        const uint32_t baseA = __ldg(cs + (c + rel - reach)) + 15u;
        const uint32_t baseB = __ldg(cs + (c - rel - reach)) + 31u;
        while (w) {
            const uint32_t b = top_bit(w);
            w ^= bit_at(b);
            const uint32_t j = (b >= 16u ? baseB : baseA) - b;
            fl.eval(__ldg(fa + j), __ldg(fb + j));
        }
    }
#pragma unroll
    for (int off = 1; off < L; off <<= 1) {
        fl.fpxy.x += __shfl_xor_sync(peers, fl.fpxy.x, off); fl.fpxy.y += __shfl_xor_sync(peers, fl.fpxy.y, off);
        fl.fpz += __shfl_xor_sync(peers, fl.fpz, off);
        fl.fvxy.x += __shfl_xor_sync(peers, fl.fvxy.x, off); fl.fvxy.y += __shfl_xor_sync(peers, fl.fvxy.y, off);
        fl.fvz += __shfl_xor_sync(peers, fl.fvz, off);
        ovf |= __shfl_xor_sync(peers, ovf, off);
    }
    if (s != 0) return;
    ForceAccum f = fl.result(a.k);
    if (ovf) {
        // some column of this particle holds more candidates than its mask has bits: the candidates beyond the mask are
        // walked with the exact radius test (like k_force_mask16)
        const float r2 = a.k.r2;
        auto rest = [&](uint32_t cell, uint32_t reach) {
            const uint32_t b = __ldg(cs + (cell - reach)), e = __ldg(cs + (cell + reach + 1u));
            for (uint32_t j = b + (uint32_t)kMaskBits; j < e; ++j) {
                const float4 pj = __ldg(&a.posm[j]);
                const float rx = __fsub_rn(pi.x, pj.x), ry = __fsub_rn(pi.y, pj.y), rz = __fsub_rn(pi.z, pj.z);
                const float d2 = dist2_exact(rx, ry, rz);
                if (d2 <= r2) {
                    const float4 qa = __ldg(fa + j), qb = __ldg(fb + j);
                    force_pair_fast(a.k, f, rx, ry, rz, d2, qb.x - vi.x, qb.y - vi.y, qb.z - vi.z, P_i, qa.w, qb.w);
                }
            }
        };
        for (int g = 0; g < ng; ++g) {
            const uint32_t rel = (uint32_t)(tab.d0[g] * e12 + tab.d1[g] * e2), reach = (uint32_t)tab.reach[g];
            rest(c + rel, reach);
            rest(c - rel, reach);
        }
        rest(c, (uint32_t)tab.reach[ng]);
    }
    a.acc[i] = accel_fast(a.k, f, pi.w);
}

template <int L>
int launch_density_split_l(const PairArgs& a, cudaStream_t st) {
    const unsigned nb = (unsigned)((a.n * L + kThreadsP - 1) / kThreadsP);
    const bool slab = a.slab_axis >= 0;
    const bool trunc = !(a.k.r2 >= 4.0f * a.k.h_sq);   // neighbor_search_radius < 2 h (see k_density_mask16)
#define SPHB_LAUNCH_D(RR)                                                                        \
    if (slab) { if (trunc) k_density_split<true, RR, true, L><<<nb, kThreadsP, 0, st>>>(a);      \
                else k_density_split<true, RR, false, L><<<nb, kThreadsP, 0, st>>>(a); }         \
    else { if (trunc) k_density_split<false, RR, true, L><<<nb, kThreadsP, 0, st>>>(a);          \
           else k_density_split<false, RR, false, L><<<nb, kThreadsP, 0, st>>>(a); }
    switch (a.walk_radius) {
        case 4: SPHB_LAUNCH_D(4); break;
        case 5: SPHB_LAUNCH_D(5); break;
        default: SPHB_LAUNCH_D(6); break;
    }
#undef SPHB_LAUNCH_D
    return 1;
}

template <int L>
int launch_force_split_l(const PairArgs& a, cudaStream_t st) {
    const unsigned nb = (unsigned)((a.n * L + kThreadsP - 1) / kThreadsP);
    const bool slab = a.slab_axis >= 0;
#define SPHB_LAUNCH_F(RR)                                                    \
    if (slab) k_force_split<true, RR, L><<<nb, kThreadsP, 0, st>>>(a);       \
    else k_force_split<false, RR, L><<<nb, kThreadsP, 0, st>>>(a)
    switch (a.walk_radius) {
        case 4: SPHB_LAUNCH_F(4); break;
        case 5: SPHB_LAUNCH_F(5); break;
        default: SPHB_LAUNCH_F(6); break;
    }
#undef SPHB_LAUNCH_F
    return 1;
}

}  // namespace

// a.lanes in {2, 4, 8}; walk radius >= 4
int launch_density_split(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    switch (a.lanes) {
        case 2: return launch_density_split_l<2>(a, st);
        case 8: return launch_density_split_l<8>(a, st);
        default: return launch_density_split_l<4>(a, st);
    }
}

int launch_force_split(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    switch (a.lanes) {
        case 2: return launch_force_split_l<2>(a, st);
        case 8: return launch_force_split_l<8>(a, st);
        default: return launch_force_split_l<4>(a, st);
    }
}

}  // namespace sphb
