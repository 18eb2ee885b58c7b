// Pair kernels, variant 2 on COARSE refined grids (walk radius R = 2, 3; 64-bit column masks).  The default fast path
// (R >= 4, ~1 particle per cell, 16-bit masks) lives in pair_mask.cu; this file is the round-1 kernel pair, kept for the
// grids the step falls back to when a refine-4 cell table would be too large (api.cu: sphb_step).
//
// The density pass hands its accepted-neighbour sets to the force pass as per-column bitmasks, so the radius test
// runs ONCE per step instead of twice.
//
// Why: both pair passes are instruction-issue bound (profiles/: 85 % / 74 % issue-active, < 10 % DRAM).  The
// tested walk of variant 0 examined ~1 000 candidates per particle and pass to find ~265 neighbours, paid ~12
// warp-instructions per test and entered the divergent "accepted" branch whenever ANY lane accepted.  Here
//   * the grid is refined to cells of nsr / R (default R = 4: ~1 particle per cell, so all lanes of a warp — 32
//     consecutive particles along a cell column — see the same relative neighbourhood) and only the cells of a
//     static spherical stencil are visited (613 of 9^3 at R = 4, constant tables, no per-lane divergence);
//   * k_density_mask walks the (2R+1)^2 cell columns (each ONE contiguous run of the sorted arrays — the
//     fast-mode layout uses monotone cell ranks and guard cells, see GridDesc), evaluates the density, and
//     records per column which candidates passed the reference's exact radius test (spatial_hash.h:70-73) as a
//     32-bit (R = 4) or 64-bit (R = 2, 3) mask.  Masks are stored column-major (masks[col][slot]) so that a warp
//     writes one fully coalesced row per column — the per-lane scattered stores that sank the neighbour-LIST
//     hand-off do not occur;
//   * k_force_mask reads the masks of its particle back (coalesced) and evaluates exactly the set bits — no
//     distance test, no rejected candidates: FLO -> slot -> one 32-byte LDG.E.256 record -> 31 flops.  Columns are
//     consumed as mirror pairs with per-lane flat bit streams, which keeps the warp's lanes balanced when the
//     particles are disordered.
// A column holding more candidates than its mask has bits (collapsed states, coincident wall layers) sets the
// particle's overflow flag; the force pass then walks the candidates BEYOND the mask of each column with the tested
// loop of variant 0, so results never depend on the mask capacity.
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
#ifndef SPHB_DMASK_UNROLL
#define SPHB_DMASK_UNROLL 4
#endif
#ifndef SPHB_FMASK_MINBLOCKS
#define SPHB_FMASK_MINBLOCKS 12   // <= 40 registers: the force pass is latency-sensitive, 48 warps/SM beat 40 (0.68 -> 0.61 ms)
#endif
#ifndef SPHB_MASK_STREAMING
#define SPHB_MASK_STREAMING 1
#endif
#define SPHB_PRAGMA(x) _Pragma(#x)
#define SPHB_UNROLL_N(n) SPHB_PRAGMA(unroll n)
#ifndef SPHB_DENSITY_F32X2
#define SPHB_DENSITY_F32X2 1
#endif
#ifndef SPHB_PIN_CONSTANTS
#define SPHB_PIN_CONSTANTS 1
#endif


// Calls body(col, valid, b, e) for the (2R+1)^2 columns around cell `center` in walk order; [b, e) is the slot run of
// the column's cells inside the spherical stencil (monotone ranks: always one contiguous run).
template <int R, typename Body>
__device__ __forceinline__ void walk_columns(const GridDesc& g, const uint32_t* __restrict__ cell_start, uint32_t center,
                                             Body&& body) {
    const uint32_t e2 = (uint32_t)g.ext[2], e12 = (uint32_t)g.ext[1] * e2;
    uint32_t row = center - (uint32_t)R * e12 - (uint32_t)R * e2;   // cell (c0 - R, c1 - R, c2)
    int col = 0;
#pragma unroll 1
    for (int d0 = -R; d0 <= R; ++d0, row += e12) {
        uint32_t idx = row;
#pragma unroll 1
        for (int d1 = -R; d1 <= R; ++d1, ++col, idx += e2) {
            const int reach = column_reach<R>(col);
            if (reach >= 0) body(col, true, __ldg(&cell_start[idx - reach]), __ldg(&cell_start[idx + reach + 1]));
            else body(col, false, 0u, 0u);
        }
    }
}

#ifndef SPHB_DMASK_MINBLOCKS
#define SPHB_DMASK_MINBLOCKS 1
#endif
// Pins a loop-invariant value in a register: ptxas otherwise re-reads kernel parameters from the constant bank inside
// the pair loops (one issue slot per use in kernels that are issue-bound).
#if SPHB_PIN_CONSTANTS
__device__ __forceinline__ float pin(float v) { asm volatile("" : "+f"(v)); return v; }
template <typename T> __device__ __forceinline__ T* pin(T* p) { asm volatile("" : "+l"(p)); return p; }
#else
__device__ __forceinline__ float pin(float v) { return v; }
template <typename T> __device__ __forceinline__ T* pin(T* p) { return p; }
#endif

// mask storage: W = 1 -> one uint32 per (column, particle), W = 2 -> one uint2
template <int W> struct MaskStore;
template <> struct MaskStore<1> {
    // masks are written once and read once: streaming (evict-first) accesses keep them from displacing the particle
    // records that the pair loops re-read from L1 / L2
#if SPHB_MASK_STREAMING
    static __device__ __forceinline__ void put(void* base, size_t idx, uint32_t lo, uint32_t) { __stcs(static_cast<uint32_t*>(base) + idx, lo); }
    static __device__ __forceinline__ uint2 get(const void* base, size_t idx) { return make_uint2(__ldcs(static_cast<const uint32_t*>(base) + idx), 0u); }
#else
    static __device__ __forceinline__ void put(void* base, size_t idx, uint32_t lo, uint32_t) { static_cast<uint32_t*>(base)[idx] = lo; }
    static __device__ __forceinline__ uint2 get(const void* base, size_t idx) { return make_uint2(static_cast<const uint32_t*>(base)[idx], 0u); }
#endif
};
template <> struct MaskStore<2> {
#if SPHB_MASK_STREAMING
    static __device__ __forceinline__ void put(void* base, size_t idx, uint32_t lo, uint32_t hi) { __stcs(static_cast<uint2*>(base) + idx, make_uint2(lo, hi)); }
    static __device__ __forceinline__ uint2 get(const void* base, size_t idx) { return __ldcs(static_cast<const uint2*>(base) + idx); }
#else
    static __device__ __forceinline__ void put(void* base, size_t idx, uint32_t lo, uint32_t hi) { static_cast<uint2*>(base)[idx] = make_uint2(lo, hi); }
    static __device__ __forceinline__ uint2 get(const void* base, size_t idx) { return static_cast<const uint2*>(base)[idx]; }
#endif
};

template <bool SLAB, int R, int W>
__global__ void __launch_bounds__(kThreads, SPHB_DMASK_MINBLOCKS) k_density_mask(PairArgs a) {
    constexpr int kMaskCols = (2 * R + 1) * (2 * R + 1);
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    unsigned count = 0;
    if (i < a.n) {
        const float4 pi = a.posm[i];
        if (!SLAB || wants_density(a, pi)) {
            const uint32_t c = center_cell(a.grid, pi);
            const float r2 = pin(a.k.r2);
            const float inv_h = pin(a.k.inv_h);
            const float4* __restrict__ posm = pin(a.posm);
            float rho = 0.0f;   // the self pair (d2 = 0) stays in the loop: the polynomial gives sigma * 4/6 there
            unsigned ovf = 0;
            const size_t stride = a.mask_stride;
            const float2 pxy = make_float2(pi.x, pi.y);
            const float nz = pin(a.k.neg_zero);
            const float2 nz2 = make_float2(nz, nz);
            // test + density contribution of slot j; returns whether j is a neighbour (exact reference test)
            // exact reference radius test of slot j against this particle: returns d2, pj
            auto dist2 = [&](const float4& pj) -> float {
#if SPHB_DENSITY_F32X2
                // (x, y) of a float4 load sit in an aligned register pair: one FADD2 + one FFMA2 (exact squares as
                // fma(d, d, -0), see pair.cu) replace two FADDs + two FMULs; every rounding is the reference's
                const float2 dxy = __fadd2_rn(pxy, make_float2(-pj.x, -pj.y));
                const float2 sq = __ffma2_rn(dxy, dxy, nz2);
                const float dz = __fsub_rn(pi.z, pj.z);
                return __fadd_rn(__fadd_rn(sq.x, sq.y), __fmul_rn(dz, dz));
#else
                return dist2_exact(__fsub_rn(pi.x, pj.x), __fsub_rn(pi.y, pj.y), __fsub_rn(pi.z, pj.z));
#endif
            };
            auto add = [&](float d2, float m) {
                const float q = fast_sqrt(d2) * inv_h;
                const float t2 = fmaxf(2.0f - q, 0.0f), t1 = fmaxf(1.0f - q, 0.0f);
                rho += m * (t2 * t2 * t2 - 4.0f * (t1 * t1 * t1));
            };
            // test + density contribution of slot j; returns whether j is a neighbour
            auto visit = [&](uint32_t j) -> bool {
                const float4 pj = __ldg(&posm[j]);
                const float d2 = dist2(pj);
                const bool in = d2 <= r2;
                if (in) add(d2, pj.w);
                return in;
            };
            walk_columns<R>(a.grid, a.cell_start, c, [&](int col, bool valid, uint32_t b, uint32_t e) {
                uint32_t mlo = 0, mhi = 0;
                if (valid) {
                    uint32_t j = b;
                    const uint32_t e1 = min(e, b + 32u);
                    uint32_t bit = 1u;
SPHB_UNROLL_N(SPHB_DMASK_UNROLL)
                    for (; j < e1; ++j, bit += bit)
                        if (visit(j)) mlo |= bit;
                    if (j < e) {
                        if (W == 2) {
                            const uint32_t e2 = min(e, b + 64u);
                            bit = 1u;
#pragma unroll 4
                            for (; j < e2; ++j, bit += bit)
                                if (visit(j)) mhi |= bit;
                        }
                        if (j < e) {   // more than 32 W candidates in this column: no mask for them
                            ovf = 1u;
                            for (; j < e; ++j)
                                if (visit(j)) ++count;
                        }
                    }
                    count += __popc(mlo) + __popc(mhi);
                }
                MaskStore<W>::put(a.masks, (size_t)col * stride + i, mlo, mhi);
            });
            MaskStore<W>::put(a.masks, (size_t)kMaskCols * stride + i, ovf, count);
            rho *= a.k.sigma * (1.0f / 6.0f);
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

#ifndef SPHB_FORCE_PIPE
#define SPHB_FORCE_PIPE 0
#endif

template <bool SLAB, int R, int W>
__global__ void __launch_bounds__(kThreads, SPHB_FMASK_MINBLOCKS) k_force_mask(PairArgs a) {
    constexpr int kMaskCols = (2 * R + 1) * (2 * R + 1);
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= a.n) return;
    const float4 vi = a.velid[i];
    if (SLAB && is_ghost(vi)) return;   // slab mode: halo copies are never advanced here
    const float4 pi = a.posm[i];
    const float P_i = a.rho_p[i].y;
    const uint32_t c = center_cell(a.grid, pi);
    const GridDesc& g = a.grid;
    ForceAccum f = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const size_t stride = a.mask_stride;
    const unsigned ovf = MaskStore<W>::get(a.masks, (size_t)kMaskCols * stride + i).x;
    // pair j -> i without a distance test (j was accepted by the density pass)
    const float inv_h = pin(a.k.inv_h);
    const ForceRec* __restrict__ fab = pin(a.fab);
    // force_pair_fast (pair_math.cuh) with the per-pair constant factors sigma/h and 2 mu sigma/h^2 taken out of the
    // sums (applied once per particle below).  1/len uses max(d2, 1e-30): coincident particles (d2 = 0) get q = 0 and
    // dW/dq(0) = 0, hence no pressure term, exactly like the reference's r_len < 1e-6 guard (sph_engine.cpp:403); a
    // distinct pair closer than 1e-6 contributes |dW/dq| <= 2e-6 / h instead of nothing — far below the fast-mode gates.
    auto eval = [&](const ForceRec& q) {
        const float rx = pi.x - q.x, ry = pi.y - q.y, rz = pi.z - q.z;
        const float d2 = rx * rx + ry * ry + rz * rz;
        const float inv_len = fast_rsqrt(fmaxf(d2, 1e-30f));
        const float qq = (d2 * inv_len) * inv_h;
        const float t2 = fmaxf(2.0f - qq, 0.0f), t1 = fmaxf(1.0f - qq, 0.0f);
        const float gq = 2.0f * (t1 * t1) - 0.5f * (t2 * t2);      // dW/dq / sigma
        const float lq = t2 - 4.0f * t1;                            // d2W/dq2 / sigma
        const float cp = (q.A * P_i + q.B) * (gq * inv_len);
        f.px -= cp * rx; f.py -= cp * ry; f.pz -= cp * rz;
        const float cv = q.A * lq;
        f.vx += cv * (q.vx - vi.x); f.vy += cv * (q.vy - vi.y); f.vz += cv * (q.vz - vi.z);
    };
    {
        // The (2R+1)^2 columns are consumed as groups {column k, its point mirror 24 - k}: a lane close to one side of
        // its cell has many neighbours in the columns on that side and few in the mirrored ones, so the SUM over a
        // mirror pair is nearly the same for all lanes of a warp.  Inside a group every lane pops its own bits as
        // one flat stream (column k, then its mirror), so the warp runs max-over-lanes(sum) iterations per group:
        // at R = 2 ~300 per particle instead of ~450 with one lock-step loop per mask word (lattice, h = 2 dx).
        const uint32_t e2 = (uint32_t)g.ext[2], e12 = (uint32_t)g.ext[1] * e2;
        uint32_t off = (uint32_t)R * e12 + (uint32_t)R * e2;   // column k is at center - off, its mirror at center + off
        int d1 = -R;
        size_t ia = i, ib = (size_t)(kMaskCols - 1) * stride + i;   // mask rows of column k and of its mirror
#pragma unroll 1
        for (int k = 0; k <= kMaskCols / 2; ++k) {
            const uint2 mA = MaskStore<W>::get(a.masks, ia);
            uint2 mB = make_uint2(0u, 0u);
            if (k < kMaskCols / 2) mB = MaskStore<W>::get(a.masks, ib);
            ia += stride; ib -= stride;
            // a column and its mirror have the same reach in the spherical stencil; masks of columns outside the
            // stencil are zero (written by the density pass) and their bases are never used
            const int reach = max(column_reach<R>(k), 0);
            const uint32_t bA = __ldg(&a.cell_start[c - off - reach]);
            const uint32_t bB = __ldg(&a.cell_start[c + off - reach]);
            off -= e2;
            if (++d1 > R) { d1 = -R; off -= e12 - (uint32_t)(2 * R + 1) * e2; }
            uint32_t lo = mA.x, hi = mA.y, base = bA;
            uint32_t lo2 = mB.x, hi2 = mB.y;
            if ((lo | hi) == 0u) { lo = lo2; hi = hi2; base = bB; lo2 = 0u; hi2 = 0u; }
            // pops the highest set bit of hi:lo and returns its slot (W == 1: hi is identically 0 and folds away)
            auto pop = [&]() -> uint32_t {
                uint32_t j;
                if (W == 2) {
                    const bool up = hi != 0u;
                    uint32_t w = up ? hi : lo;
                    const int b = 31 - __clz(w);
                    w ^= 1u << b;
                    if (up) hi = w; else lo = w;
                    j = base + (uint32_t)b + (up ? 32u : 0u);
                    if ((lo | hi) == 0u) { lo = lo2; hi = hi2; base = bB; lo2 = 0u; hi2 = 0u; }
                } else {
                    const int b = 31 - __clz(lo);
                    lo ^= 1u << b;
                    j = base + (uint32_t)b;
                    if (lo == 0u) { lo = lo2; base = bB; lo2 = 0u; }
                }
                return j;
            };
#if SPHB_FORCE_PIPE
            bool have = (lo | hi) != 0u;
            ForceRec nxt;
            if (have) nxt = load_rec(fab + pop());
            while (have) {
                const ForceRec cur = nxt;
                have = (lo | hi) != 0u;
                if (have) nxt = load_rec(fab + pop());
                eval(cur);
            }
#else
            while (lo | hi) eval(load_rec(fab + pop()));
#endif
        }
    }
    f.px *= a.k.sig_h; f.py *= a.k.sig_h; f.pz *= a.k.sig_h;
    {
        const float cvis = 2.0f * a.k.viscosity * a.k.sig_h2;
        f.vx *= cvis; f.vy *= cvis; f.vz *= cvis;
    }
    if (ovf) {
        // some column of this particle holds more candidates than its mask has bits (collapsed states, coincident
        // wall layers): the candidates beyond the mask are walked with the exact radius test, like variant 0
        const float r2 = a.k.r2;
        walk_columns<R>(a.grid, a.cell_start, c, [&](int col, bool valid, uint32_t b, uint32_t e) {
            for (uint32_t j = b + 32u * W; j < e; ++j) {
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
        });
    }
    a.acc[i] = accel_fast(a.k, f, pi.w);
}

}  // namespace

int launch_density_mask_wide(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    const unsigned nb = (unsigned)((a.n + kThreads - 1) / kThreads);
    const bool slab = a.slab_axis >= 0;
#define SPHB_LAUNCH_D(RR, WW)                                                 \
    if (slab) k_density_mask<true, RR, WW><<<nb, kThreads, 0, st>>>(a);       \
    else k_density_mask<false, RR, WW><<<nb, kThreads, 0, st>>>(a)
    if (a.walk_radius == 2) { SPHB_LAUNCH_D(2, 2); } else { SPHB_LAUNCH_D(3, 2); }
#undef SPHB_LAUNCH_D
    return 1;
}

int launch_force_mask_wide(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    const unsigned nb = (unsigned)((a.n + kThreads - 1) / kThreads);
    const bool slab = a.slab_axis >= 0;
#define SPHB_LAUNCH_F(RR, WW)                                                 \
    if (slab) k_force_mask<true, RR, WW><<<nb, kThreads, 0, st>>>(a);         \
    else k_force_mask<false, RR, WW><<<nb, kThreads, 0, st>>>(a)
    if (a.walk_radius == 2) { SPHB_LAUNCH_F(2, 2); } else { SPHB_LAUNCH_F(3, 2); }
#undef SPHB_LAUNCH_F
    return 1;
}

}  // namespace sphb
