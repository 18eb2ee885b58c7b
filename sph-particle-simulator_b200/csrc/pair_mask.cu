// Pair kernels of the default fast path (fast math, refined grid with R = 4..6 cells per search radius, ~1 particle
// per cell): the density pass finds every particle's neighbour set ONCE per step and hands it to the force pass as
// 16-bit masks per cell column, two columns (a column and its point mirror) per 32-bit word.
//
// Round-2 redesign of the round-1 kernels (now pair_mask_wide.cu), driven by profiles/r1d_hot_regions.md: the old
// density pass spent 28 warp-instructions per candidate (predicated scalar evaluation, 65 instructions of bookkeeping
// per column, 1-3 candidate remainder loops).  Here
//   * candidates are processed TWO per iteration: the (x, y) pair of each float4 record and the pairs formed across
//     the two candidates (dz, the partial sums, everything of the B-spline evaluation) run on the packed f32x2 pipe
//     (FADD2 / FFMA2 / FMUL2), an odd tail is the same iteration with its second candidate pushed out of range;
//   * the radius test is still the reference's exact one (spatial_hash.h:70-73: every product and sum rounded
//     separately); its outcome is taken from the SIGN of d2 - nextafter(r2) and shifted into the mask with one
//     funnel shift — no compare, no predicate, no select;
//   * the density contribution is evaluated unconditionally in compact-support form ((2-q)+^3 - 4 (1-q)+^3 is exactly 0
//     beyond the support), so there is no divergent "accepted" branch and no remainder loop;
//   * a column and its mirror share one mask word, one row store, one row load in the force pass (half the mask
//     traffic of round 1: 39 instead of 82 rows at R = 4), and the force pass pops both columns as ONE flat 32-bit
//     bit stream, which keeps the lanes of a warp balanced (a lane near one face of its cell has many neighbours on
//     that side and few on the mirrored side);
//   * columns holding more than 16 candidates (collapsed states, coincident wall layers) set the particle's overflow
//     flag; both passes walk the candidates beyond the mask with the tested scalar loop, so results never depend on
//     the mask capacity.
//
// Reference: SPHEngine::update_neighbor_lists' query + compute_densities + compute_pressures + compute_forces
// (src/sph_engine.cpp:335-353, 203-244); the bitmask is this design's stand-in for neighbor_lists_[i].
#include "pair_stencil.cuh"

namespace sphb {

namespace {

#ifndef SPHB_MASK_THREADS
#define SPHB_MASK_THREADS 128
#endif
constexpr int kThreads = SPHB_MASK_THREADS;
#ifndef SPHB_DMASK_MINBLOCKS
#define SPHB_DMASK_MINBLOCKS 1
#endif
#ifndef SPHB_FMASK_MINBLOCKS
#define SPHB_FMASK_MINBLOCKS 12   // <= 40 registers: the force pass is latency-sensitive, 48 warps/SM beat 40
#endif
#ifndef SPHB_DMASK_GROUP_UNROLL
#define SPHB_DMASK_GROUP_UNROLL 2
#endif
#ifndef SPHB_FORCE_PIPE
#define SPHB_FORCE_PIPE 0
#endif
#define SPHB_PRAGMA(x) _Pragma(#x)
#define SPHB_UNROLL_N(n) SPHB_PRAGMA(unroll n)

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// Pins a loop-invariant value in a register: ptxas otherwise re-reads kernel parameters from the constant bank inside
// the pair loops (one issue slot per use in kernels that are issue-bound).
#ifndef SPHB_PIN_CONSTANTS
#define SPHB_PIN_CONSTANTS 1
#endif
#if SPHB_PIN_CONSTANTS
__device__ __forceinline__ float pin(float v) { asm volatile("" : "+f"(v)); return v; }
__device__ __forceinline__ uint32_t pin(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
template <typename T> __device__ __forceinline__ T* pin(T* p) { asm volatile("" : "+l"(p)); return p; }
#else
__device__ __forceinline__ float pin(float v) { return v; }
__device__ __forceinline__ uint32_t pin(uint32_t v) { return v; }
template <typename T> __device__ __forceinline__ T* pin(T* p) { return p; }
#endif
// index of the highest set bit (FLO)
__device__ __forceinline__ uint32_t top_bit(uint32_t w) {
    uint32_t b;
    asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(w));
    return b;
}
// bit `pos` as a mask (one BMSK instead of materialising a constant and shifting it)
__device__ __forceinline__ uint32_t bit_at(uint32_t pos) {
    uint32_t m;
    asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(m) : "r"(pos));
    return m;
}

// Density + EOS of one particle per thread, candidates read from global memory (DensityWalker, pair_stencil.cuh).
// TRUNC: the search radius cuts the kernel support short (neighbor_search_radius < 2 h).
template <bool SLAB, int R, bool TRUNC>
__global__ void __launch_bounds__(kThreads, SPHB_DMASK_MINBLOCKS) k_density_mask16(PairArgs a) {
    __shared__ int4 soff[Groups<R>::kGroups + 1];   // per group: cell offsets of {run start, run end} of column and mirror
    const size_t i_raw = (size_t)blockIdx.x * kThreads + threadIdx.x;
    // Lanes past the last particle walk as copies of it (no special cases inside the loops) and store nothing at the end
    const size_t i = i_raw < a.n ? i_raw : a.n - 1;
    const float4 pi = a.posm[i];
    const bool want = i_raw < a.n && (!SLAB || wants_density(a, pi));
    const GroupTable<R>& tab = group_table<R>();
    const int ng = tab.n;   // groups 0 .. ng - 1 are mirror pairs, group ng is the centre column
    {
        const int e2 = a.grid.ext[2], e12 = a.grid.ext[1] * e2;
        for (int g = threadIdx.x; g <= ng; g += kThreads) {
            const int rel = tab.d0[g] * e12 + tab.d1[g] * e2, reach = tab.reach[g];
            soff[g] = make_int4(rel - reach, rel + reach + 1, -rel - reach, -rel + reach + 1);
        }
    }
    __syncthreads();
    unsigned count = 0;
    if (want) {
        const uint32_t* __restrict__ csc = a.cell_start + center_cell(a.grid, pi);
        const float4* __restrict__ posm = pin(a.posm);
        DensityWalker<TRUNC> dw;
        dw.init(pi, a.k, (uint32_t)(a.n >> 62));
        // Walk over the mirror-pair groups, software-pipelined ACROSS groups: the run bounds of group g + 1 are loaded
        // while group g is processed (without this every column starts with two dependent round trips, cell table ->
        // records)
        auto bounds = [&](int g, uint32_t& b0, uint32_t& e0, uint32_t& b1, uint32_t& e1) {
            const int4 o = soff[g];
            b0 = __ldg(csc + o.x); e0 = __ldg(csc + o.y);
            b1 = __ldg(csc + o.z); e1 = __ldg(csc + o.w);
        };
        uint32_t bA, eA, bB, eB;
        bounds(0, bA, eA, bB, eB);
        uint32_t* __restrict__ mrow = static_cast<uint32_t*>(a.masks) + i;
        const size_t stride = a.mask_stride;
SPHB_UNROLL_N(SPHB_DMASK_GROUP_UNROLL)
        for (int g = 0; g < ng; ++g) {
            uint32_t nbA, neA, nbB, neB;
            bounds(g + 1, nbA, neA, nbB, neB);   // entry ng is the centre column (both halves the same run)
            uint32_t word = dw.column(posm, bA, eA, 0u, posm, pi, a.k);
            word |= dw.column(posm, bB, eB, 0u, posm, pi, a.k) << 16;
            count += __popc(word);
            __stcs(mrow, word);
            mrow += stride;
            bA = nbA; eA = neA; bB = nbB; eB = neB;
        }
        {
            const uint32_t word = dw.column(posm, bA, eA, 0u, posm, pi, a.k);
            count += __popc(word);
            __stcs(mrow, word | (dw.ovf << 31));   // bit 31 of the centre word (its column only uses the low half): the overflow flag
        }
        count += dw.extra;

        const float rho = dw.density(a.k);
        const float P = a.k.gas_constant * (rho - a.k.rest_density);
        a.rho_p[i] = make_float2(rho, P);
        const float4 v = a.velid[i];
        const float A = pi.w / (2.0f * rho);
        // force-pass records in two arrays of 16-byte halves (the layout the staged kernels gather without bank conflicts)
        a.fa[i] = make_float4(pi.x, pi.y, pi.z, A);
        a.fb[i] = make_float4(v.x, v.y, v.z, A * P);
        if (a.nbr_count) a.nbr_count[i] = count;
    }
    count = __reduce_max_sync(0xffffffffu, count);
    if ((threadIdx.x & 31) == 0 && count > *(volatile unsigned int*)&a.sc->max_neighbors) atomicMax(&a.sc->max_neighbors, count);
}

template <bool SLAB, int R>
__global__ void __launch_bounds__(kThreads, SPHB_FMASK_MINBLOCKS) k_force_mask16(PairArgs a) {
    constexpr int kGroups = Groups<R>::kGroups;
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= a.n) return;
    const float4 vi = a.velid[i];
    if (SLAB && is_ghost(vi)) return;   // slab mode: halo copies are never advanced here
    const float4 pi = a.posm[i];
    const float P_i = a.rho_p[i].y;
    const uint32_t c = center_cell(a.grid, pi);
    const uint32_t* __restrict__ cs = a.cell_start;
    const float4* __restrict__ fa = pin(a.fa);
    const float4* __restrict__ fb = pin(a.fb);
    auto load2 = [&](uint32_t j) -> ForceRec {
        const float4 qa = __ldg(fa + j), qb = __ldg(fb + j);
        ForceRec r;
        r.x = qa.x; r.y = qa.y; r.z = qa.z; r.A = qa.w; r.vx = qb.x; r.vy = qb.y; r.vz = qb.z; r.B = qb.w;
        return r;
    };
    ForceLane fl;
    fl.init(pi, vi, P_i, a.k, (uint32_t)(a.n >> 62));
    auto eval = [&](const ForceRec& q) { fl.eval(make_float4(q.x, q.y, q.z, q.A), make_float4(q.vx, q.vy, q.vz, q.B)); };

    const int e2 = a.grid.ext[2], e12 = a.grid.ext[1] * e2;
    uint32_t rel = (uint32_t)(R * e12 + R * e2);
    int d1 = -R;
    const uint32_t* __restrict__ mrow = static_cast<const uint32_t*>(a.masks) + i;
    const size_t stride = a.mask_stride;
    unsigned ovf = 0;
#pragma unroll 1
    for (int g = 0; g <= kGroups; ++g) {
        const int reach = column_reach<R>(g);
        if (reach >= 0) {   // warp-uniform; a column and its mirror have the same reach
            uint32_t w = __ldcs(mrow);
            mrow += stride;
            if (g == kGroups) { ovf = w >> 31; w &= 0xFFFFu; }
            // candidate q of column g is bit 15 - q, of its mirror bit 31 - q: slot = base - bit
            const uint32_t baseA = __ldg(cs + (c - rel - (uint32_t)reach)) + 15u;
            const uint32_t baseB = __ldg(cs + (c + rel - (uint32_t)reach)) + 31u;
            // every lane pops the bits of both columns as one flat stream: the warp runs max-over-lanes(popcount)
            // iterations per group, and that sum over a mirror pair is nearly lane-independent
#if SPHB_FORCE_PIPE
            ForceRec nxt;
            bool have = w != 0u;
            if (have) {
                const uint32_t b = top_bit(w);
                w ^= bit_at(b);
                nxt = load2((b >= 16u ? baseB : baseA) - b);
            }
            while (have) {
                const ForceRec cur = nxt;
                have = w != 0u;
                if (have) {
                    const uint32_t b = top_bit(w);
                    w ^= bit_at(b);
                    nxt = load2((b >= 16u ? baseB : baseA) - b);
                }
                eval(cur);
            }
#else
            while (w) {
                const uint32_t b = top_bit(w);
                w ^= bit_at(b);
                eval(load2((b >= 16u ? baseB : baseA) - b));
            }
#endif
        }
        rel -= (uint32_t)e2;
        if (++d1 > R) { d1 = -R; rel -= (uint32_t)(e12 - (2 * R + 1) * e2); }
    }
    ForceAccum f = fl.result(a.k);
    if (ovf) {
        // some column of this particle holds more candidates than its mask has bits: the candidates beyond the mask
        // are walked with the exact radius test
        const float r2 = a.k.r2;
        auto rest = [&](uint32_t cell, int reach) {
            const uint32_t b = __ldg(cs + (cell - (uint32_t)reach)), e = __ldg(cs + (cell + (uint32_t)reach + 1u));
            for (uint32_t j = b + (uint32_t)kMaskBits; j < e; ++j) {
                // positions from posm: in slab mode the records are only written where the density was evaluated (owned +
                // first halo layer), which covers every ACCEPTED j of an owned particle but not every candidate
                const float4 pj = __ldg(&a.posm[j]);
                const float rx = __fsub_rn(pi.x, pj.x), ry = __fsub_rn(pi.y, pj.y), rz = __fsub_rn(pi.z, pj.z);
                const float d2 = dist2_exact(rx, ry, rz);
                if (d2 <= r2) {
                    const ForceRec q = load2(j);
                    force_pair_fast(a.k, f, rx, ry, rz, d2, q.vx - vi.x, q.vy - vi.y, q.vz - vi.z, P_i, q.A, q.B);
                }
            }
        };
        uint32_t rel2 = (uint32_t)(R * e12 + R * e2);
        int dd1 = -R;
        for (int g = 0; g < kGroups; ++g) {
            const int reach = column_reach<R>(g);
            if (reach >= 0) { rest(c - rel2, reach); rest(c + rel2, reach); }
            rel2 -= (uint32_t)e2;
            if (++dd1 > R) { dd1 = -R; rel2 -= (uint32_t)(e12 - (2 * R + 1) * e2); }
        }
        rest(c, R);
    }
    a.acc[i] = accel_fast(a.k, f, pi.w);
}

template <int R>
constexpr int count_rows() {
    int rows = 1;   // centre column
    for (int g = 0; g < Groups<R>::kGroups; ++g) {
        const int n = 2 * R + 1;
        if (reach_of(R, g / n - R, g % n - R) >= 0) ++rows;
    }
    return rows;
}

}  // namespace

// Host copy of the stencil tables (the device reads the identical constexpr data from constant memory): lets the
// CPU tests check the shipped stencil against a brute-force model without a GPU.
int stencil_reach_table(int R, signed char* out) {
    constexpr ReachTables t = make_reach_tables();
    const int* src = R == 2 ? t.r2 : R == 3 ? t.r3 : R == 4 ? t.r4 : R == 5 ? t.r5 : R == 6 ? t.r6 : nullptr;
    if (!src) return -1;
    for (int k = 0; k < mask_cols(R); ++k) out[k] = (signed char)src[k];
    return mask_cols(R);
}

// bytes of neighbour-mask storage per unit of mask_stride
size_t mask_bytes_per_slot(int R) {
    switch (R) {
        case 2: return (size_t)(mask_cols(2) + 1) * sizeof(uint2);
        case 3: return (size_t)(mask_cols(3) + 1) * sizeof(uint2);
        case 4: return (size_t)count_rows<4>() * sizeof(uint32_t);
        case 5: return (size_t)count_rows<5>() * sizeof(uint32_t);
        default: return (size_t)count_rows<6>() * sizeof(uint32_t);
    }
}

int launch_density_mask(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    if (a.walk_radius < 4) return launch_density_mask_wide(a, st);
    if (a.lanes > 1) return launch_density_split(a, st);
    if (a.mode != 0) return launch_density_stage(a, st);
    const unsigned nb = (unsigned)((a.n + kThreads - 1) / kThreads);
    const bool slab = a.slab_axis >= 0;
    // neighbor_search_radius < 2 h: the kernel support is truncated by the search radius (see k_density_mask16)
    const bool trunc = !(a.k.r2 >= 4.0f * a.k.h_sq);
#define SPHB_LAUNCH_D(RR)                                                                     \
    if (slab) { if (trunc) k_density_mask16<true, RR, true><<<nb, kThreads, 0, st>>>(a);      \
                else k_density_mask16<true, RR, false><<<nb, kThreads, 0, st>>>(a); }         \
    else { if (trunc) k_density_mask16<false, RR, true><<<nb, kThreads, 0, st>>>(a);          \
           else k_density_mask16<false, RR, false><<<nb, kThreads, 0, st>>>(a); }
    switch (a.walk_radius) {
        case 4: SPHB_LAUNCH_D(4); break;
        case 5: SPHB_LAUNCH_D(5); break;
        default: SPHB_LAUNCH_D(6); break;
    }
#undef SPHB_LAUNCH_D
    return 1;
}

int launch_force_mask(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    if (a.walk_radius < 4) return launch_force_mask_wide(a, st);
    if (a.lanes > 1) return launch_force_split(a, st);
    if (a.mode == 1) return launch_force_stage(a, st);
    const unsigned nb = (unsigned)((a.n + kThreads - 1) / kThreads);
    const bool slab = a.slab_axis >= 0;
#define SPHB_LAUNCH_F(RR)                                                 \
    if (slab) k_force_mask16<true, RR><<<nb, kThreads, 0, st>>>(a);       \
    else k_force_mask16<false, RR><<<nb, kThreads, 0, st>>>(a)
    switch (a.walk_radius) {
        case 4: SPHB_LAUNCH_F(4); break;
        case 5: SPHB_LAUNCH_F(5); break;
        default: SPHB_LAUNCH_F(6); break;
    }
#undef SPHB_LAUNCH_F
    return 1;
}

}  // namespace sphb
