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
#ifndef SPHB_DMASK_UNROLL
#define SPHB_DMASK_UNROLL 1
#endif
#ifndef SPHB_FORCE_PIPE
#define SPHB_FORCE_PIPE 0
#endif
#ifndef SPHB_DMASK_PREFETCH
#define SPHB_DMASK_PREFETCH 0
#endif
#ifndef SPHB_EXPERIMENT
#define SPHB_EXPERIMENT 0
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
// 16-byte record load that the compiler must leave where it is written (volatile): the software-pipelined loops issue
// the loads of the NEXT candidates before the arithmetic of the current ones, and NVVM otherwise sinks them to their use
__device__ __forceinline__ float4 ldg_here(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
// bit `pos` as a mask (one BMSK instead of materialising a constant and shifting it)
__device__ __forceinline__ uint32_t bit_at(uint32_t pos) {
    uint32_t m;
    asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(m) : "r"(pos));
    return m;
}

// TRUNC: the search radius cuts the kernel support short (neighbor_search_radius < 2 h, e.g. the reference's dam-break
// example: h = 0.025, radius 0.04): candidates between the two radii must not contribute although W > 0 there, so the
// weight is additionally gated by the accept bit.  With radius >= 2 h the compact-support form alone is exact.
template <bool SLAB, int R, bool TRUNC>
__global__ void __launch_bounds__(kThreads, SPHB_DMASK_MINBLOCKS) k_density_mask16(PairArgs a) {
    constexpr int kGroups = Groups<R>::kGroups;
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    unsigned count = 0;
    if (i < a.n) {
        const float4 pi = a.posm[i];
        if (!SLAB || wants_density(a, pi)) {
            const uint32_t c = center_cell(a.grid, pi);
            const uint32_t* __restrict__ cs = pin(a.cell_start);
            const float4* __restrict__ posm = pin(a.posm);
            const float2 npxy = f2(-pi.x, -pi.y);
            const float npz = -pi.z;
            const float nzf = pin(a.k.neg_zero);
            const float2 nz2 = f2(nzf, nzf);
            const float nr2f = pin(-a.k.r2_next);
            const float2 nr2 = f2(nr2f, nr2f);
            const float ninvhf = pin(-a.k.inv_h);
            const float2 ninvh = f2(ninvhf, ninvhf);
            const float2 two2 = f2(2.0f, 2.0f);
            // (1-q)+ is carried as c4 (1-q)+ with c4 = 4^(1/3), so that W * 6 / sigma = t2^3 - t1^3
            const float2 c4 = f2(1.587401052f, 1.587401052f);
            const float2 nc4invh = f2(-1.587401052f * a.k.inv_h, -1.587401052f * a.k.inv_h);
            float rho0 = 0.0f, rho1 = 0.0f;   // the self pair (d2 = 0) stays in the loop: the polynomial gives sigma * 4/6 there
            unsigned ovf = 0;

            // squared distances of slots j, j+1 to this particle with the reference's roundings:
            // fl(fl(fl(dx dx) + fl(dy dy)) + fl(dz dz)); squares as fma(d, d, -0) (exact product; ptxas would fuse a
            // packed multiply with the following packed add, see pair.cu)
            auto dist2_pair = [&](const float4& pa, const float4& pb) -> float2 {
                const float2 da = __fadd2_rn(f2(pa.x, pa.y), npxy), db = __fadd2_rn(f2(pb.x, pb.y), npxy);
                const float2 dz = f2(__fadd_rn(pa.z, npz), __fadd_rn(pb.z, npz));
                const float2 sa = __ffma2_rn(da, da, nz2), sb = __ffma2_rn(db, db, nz2), sz = __ffma2_rn(dz, dz, nz2);
                return __fadd2_rn(f2(__fadd_rn(sa.x, sa.y), __fadd_rn(sb.x, sb.y)), sz);
            };
            // W * 6 / sigma of both candidates, exactly 0 for q >= 2
            auto weight_pair = [&](const float2& d2) -> float2 {
                const float2 s = f2(fast_sqrt(d2.x), fast_sqrt(d2.y));
                float2 t2 = __ffma2_rn(s, ninvh, two2);
                float2 t1 = __ffma2_rn(s, nc4invh, c4);
                t2.x = fmaxf(t2.x, 0.0f); t2.y = fmaxf(t2.y, 0.0f);
                t1.x = fmaxf(t1.x, 0.0f); t1.y = fmaxf(t1.y, 0.0f);
                const float2 t2c = __fmul2_rn(__fmul2_rn(t2, t2), t2);
                const float2 nt1s = __fmul2_rn(t1, f2(-t1.x, -t1.y));
                return __ffma2_rn(nt1s, t1, t2c);
            };
            // test + density contribution of the candidate pair (pa, pb); `single`: pb is not part of this run (it may
            // even be a neighbour that belongs to another column) — push it out of range
            uint32_t m;
#if SPHB_EXPERIMENT == 4
            uint32_t m2 = 0; float rho2 = 0.0f, rho3 = 0.0f; const float xshift = pin(a.k.h * 1e-3f);
#endif
            auto visit = [&](const float4& pa, const float4& pb, bool single) {
                float2 d2 = dist2_pair(pa, pb);
                if (single) d2.y = 3.0e38f;
                // accepted <=> d2 <= r2 <=> d2 - nextafter(r2) < 0: the sign bit, NaN gives 0 like the reference's compare
                const float2 t = __fadd2_rn(d2, nr2);
                m = __funnelshift_l(__float_as_uint(t.x), m, 1);
                m = __funnelshift_l(__float_as_uint(t.y), m, 1);
                float2 w = weight_pair(d2);
                if (TRUNC) {
                    w.x = t.x < 0.0f ? w.x : 0.0f;
                    w.y = t.y < 0.0f ? w.y : 0.0f;
                }
                rho0 = fmaf(pa.w, w.x, rho0);
                rho1 = fmaf(pb.w, w.y, rho1);
#if SPHB_EXPERIMENT == 4   // timing experiment: a second (fake) target on the same loads — is the pass bound by math or by L1?
                {
                    const float4 pa2 = make_float4(pa.x + xshift, pa.y, pa.z, pa.w), pb2 = make_float4(pb.x + xshift, pb.y, pb.z, pb.w);
                    float2 e2 = dist2_pair(pa2, pb2);
                    if (single) e2.y = 3.0e38f;
                    const float2 t2 = __fadd2_rn(e2, nr2);
                    m2 = __funnelshift_l(__float_as_uint(t2.x), m2, 1);
                    m2 = __funnelshift_l(__float_as_uint(t2.y), m2, 1);
                    const float2 w2 = weight_pair(e2);
                    rho2 = fmaf(pa.w, w2.x, rho2);
                    rho3 = fmaf(pb.w, w2.y, rho3);
                }
#endif
            };
            // One cell column: slots [b, e) of the sorted arrays.  Returns the 16-bit mask, candidate q at bit 15 - q.
            auto column = [&](uint32_t b, uint32_t e) -> uint32_t {
                const uint32_t end = min(e, b + (uint32_t)kMaskBits);
                m = 0;
                uint32_t j = b;
SPHB_UNROLL_N(SPHB_DMASK_UNROLL)
                for (; j < end; j += 2) {
                    // slot end may be read (as the masked second half of an odd tail): slot n is a finite sentinel
                    const float4 pa = __ldg(posm + j), pb = __ldg(posm + j + 1);
                    visit(pa, pb, j + 1 >= end);
                }
                m <<= (b + (uint32_t)kMaskBits) - j;   // j - b slots were shifted in (an even number <= 16)
                if (e > end) {   // more than 16 candidates in this column: no mask for the rest
                    ovf = 1u;
                    const float r2 = a.k.r2, inv_h = a.k.inv_h;
                    for (uint32_t u = end; u < e; ++u) {
                        const float4 pj = __ldg(posm + u);
                        const float d2 = dist2_exact(__fsub_rn(pi.x, pj.x), __fsub_rn(pi.y, pj.y), __fsub_rn(pi.z, pj.z));
                        if (d2 <= r2) {
                            ++count;
                            const float q = fast_sqrt(d2) * inv_h;
                            const float t2 = fmaxf(2.0f - q, 0.0f), t1 = fmaxf(1.0f - q, 0.0f);
                            rho0 += pj.w * (t2 * t2 * t2 - 4.0f * (t1 * t1 * t1));
                        }
                    }
                }
                return m;
            };

            // Walk over the mirror-pair groups, software-pipelined ACROSS groups: the run bounds of group g + 1 are
            // loaded while group g is processed, and once they have arrived (between its two columns) the first
            // lines of the next runs are prefetched into L1.  Without this every column starts with two dependent
            // L2 round trips (cell table -> records), ~30 % of the pass (profiles/r2_density_latency.md).
            const GroupTable<R>& tab = group_table<R>();
            const int e2 = a.grid.ext[2], e12 = a.grid.ext[1] * e2;
            uint32_t bA, eA, bB, eB;
            auto bounds = [&](int g, uint32_t& b0, uint32_t& e0, uint32_t& b1, uint32_t& e1) {
                const uint32_t rel = (uint32_t)(tab.d0[g] * e12 + tab.d1[g] * e2), reach = (uint32_t)tab.reach[g];
                b0 = __ldg(cs + (c + rel - reach)); e0 = __ldg(cs + (c + rel + reach + 1u));
                b1 = __ldg(cs + (c - rel - reach)); e1 = __ldg(cs + (c - rel + reach + 1u));
            };
            bounds(0, bA, eA, bB, eB);
            uint32_t* __restrict__ mrow = static_cast<uint32_t*>(a.masks) + i;
            const size_t stride = a.mask_stride;
            const int ngroups = tab.n;
#pragma unroll 1
            for (int g = 0; g < ngroups; ++g) {
                uint32_t nbA, neA, nbB, neB;
                bounds(g + 1, nbA, neA, nbB, neB);   // entry n is the centre column (both halves the same run)
                uint32_t word = column(bA, eA);
#if SPHB_DMASK_PREFETCH
                prefetch_l1(posm + nbA); prefetch_l1(posm + nbA + 8);
                prefetch_l1(posm + nbB); prefetch_l1(posm + nbB + 8);
#endif
                word |= column(bB, eB) << 16;
                count += __popc(word);
                __stcs(mrow, word);
                mrow += stride;
                bA = nbA; eA = neA; bB = nbB; eB = neB;
            }
            {
                const uint32_t word = column(bA, eA);
                count += __popc(word);
                __stcs(mrow, word | (ovf << 31));
            }

#if SPHB_EXPERIMENT == 4
            if (m2 == 0x12345u && rho2 + rho3 == 1.2345f) rho0 += 1.0f;   // keeps the fake target alive
#endif
            const float rho = (rho0 + rho1) * (a.k.sigma * (1.0f / 6.0f));
            const float P = a.k.gas_constant * (rho - a.k.rest_density);
            a.rho_p[i] = make_float2(rho, P);
            const float4 v = a.velid[i];
            const float A = pi.w / (2.0f * rho);
            float4* rec = reinterpret_cast<float4*>(a.fab + i);
            rec[0] = make_float4(pi.x, pi.y, pi.z, A);
            rec[1] = make_float4(v.x, v.y, v.z, A * P);
            if (a.nbr_count) a.nbr_count[i] = count;
        }
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
    const ForceRec* __restrict__ fab = pin(a.fab);
    const float2 npxy = f2(-pi.x, -pi.y), nvxy = f2(-vi.x, -vi.y);
    const float ninvh = pin(-a.k.inv_h);
    // accumulators of -F_pressure / (sigma / h) and F_viscosity / (2 mu sigma / h^2): (x, y) packed, z scalar
    float2 fpxy = f2(0.0f, 0.0f), fvxy = f2(0.0f, 0.0f);
    float fpz = 0.0f, fvz = 0.0f;
    // pair j -> i without a distance test (j was accepted by the density pass): force_pair_fast (pair_math.cuh) with
    // the per-pair constant factors taken out of the sums and r = p_j - p_i (the sign is applied once at the end).
    // 1/len uses max(d2, 1e-30): coincident particles (d2 = 0) get q = 0 and dW/dq(0) = 0, hence no pressure term,
    // exactly like the reference's r_len < 1e-6 guard (sph_engine.cpp:403); a distinct pair closer than 1e-6
    // contributes |dW/dq| <= 2e-6 / h instead of nothing — far below the fast-mode gates.
    auto eval = [&](const ForceRec& q) {
        const float2 rxy = __fadd2_rn(f2(q.x, q.y), npxy);
        const float rz = q.z - pi.z;
        const float d2 = fmaf(rz, rz, fmaf(rxy.y, rxy.y, rxy.x * rxy.x));
        const float inv_len = fast_rsqrt(fmaxf(d2, 1e-30f));
        const float t2 = fmaxf(fmaf(d2 * ninvh, inv_len, 2.0f), 0.0f);   // (2 - q)+
        const float t1 = fmaxf(t2 - 1.0f, 0.0f);                          // (1 - q)+
        const float gh = fmaf(-0.25f * t2, t2, t1 * t1);                  // dW/dq / (2 sigma) = (1-q)+^2 - (2-q)+^2 / 4
        const float lq = fmaf(-4.0f, t1, t2);                             // d2W/dq2 / sigma
        const float cp = fmaf(q.A, P_i, q.B) * (gh * inv_len);
        fpxy = __ffma2_rn(f2(cp, cp), rxy, fpxy);
        fpz = fmaf(cp, rz, fpz);
        const float cv = q.A * lq;
        const float2 uxy = __fadd2_rn(f2(q.vx, q.vy), nvxy);
        fvxy = __ffma2_rn(f2(cv, cv), uxy, fvxy);
        fvz = fmaf(cv, q.vz - vi.z, fvz);
    };

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
                nxt = load_rec(fab + ((b >= 16u ? baseB : baseA) - b));
            }
            while (have) {
                const ForceRec cur = nxt;
                have = w != 0u;
                if (have) {
                    const uint32_t b = top_bit(w);
                    w ^= bit_at(b);
                    nxt = load_rec(fab + ((b >= 16u ? baseB : baseA) - b));
                }
                eval(cur);
            }
#else
            while (w) {
                const uint32_t b = top_bit(w);
                w ^= bit_at(b);
                eval(load_rec(fab + ((b >= 16u ? baseB : baseA) - b)));
            }
#endif
        }
        rel -= (uint32_t)e2;
        if (++d1 > R) { d1 = -R; rel -= (uint32_t)(e12 - (2 * R + 1) * e2); }
    }
    ForceAccum f;
    {
        // F_p = -sum m_j term gradW with gradW along p_i - p_j = -r and gh = dW/dq / 2: the two signs cancel
        const float sp = 2.0f * a.k.sig_h;
        const float cvis = 2.0f * a.k.viscosity * a.k.sig_h2;
        f.px = sp * fpxy.x; f.py = sp * fpxy.y; f.pz = sp * fpz;
        f.vx = cvis * fvxy.x; f.vy = cvis * fvxy.y; f.vz = cvis * fvz;
    }
    if (ovf) {
        // some column of this particle holds more candidates than its mask has bits: the candidates beyond the mask
        // are walked with the exact radius test
        const float r2 = a.k.r2;
        auto rest = [&](uint32_t cell, int reach) {
            const uint32_t b = __ldg(cs + (cell - (uint32_t)reach)), e = __ldg(cs + (cell + (uint32_t)reach + 1u));
            for (uint32_t j = b + (uint32_t)kMaskBits; j < e; ++j) {
                // positions from posm: in slab mode fab is only written where the density was evaluated (owned +
                // first halo layer), which covers every ACCEPTED j of an owned particle but not every candidate
                const float4 pj = __ldg(&a.posm[j]);
                const float rx = __fsub_rn(pi.x, pj.x), ry = __fsub_rn(pi.y, pj.y), rz = __fsub_rn(pi.z, pj.z);
                const float d2 = dist2_exact(rx, ry, rz);
                if (d2 <= r2) {
                    const ForceRec q = load_rec(a.fab + j);
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
    if (a.mode == 1) return launch_density_stage(a, st);
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
