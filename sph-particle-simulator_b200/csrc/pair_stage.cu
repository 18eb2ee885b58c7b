// Staged pair kernels of the fast path (R >= 4): every WARP brings the neighbour-cell particles of its 32 targets into
// shared memory with asynchronous copies, one cell-column group ahead of their use, and its lanes traverse them from
// there (SPHB_OPT_PAIR_MODE 1).
//
// Same arithmetic, same neighbour masks and same summation order as the per-lane kernels of pair_mask.cu — results are
// bit-identical (tests/test_gpu_parity.py::test_staged_equals_global); what changes is WHERE a candidate is read from.
// In pair_mask.cu every lane fetches each of its ~600 candidates with a private LDG; ncu (profiles/
// r2b_pair_kernels_keys.txt) shows both passes waiting on those loads (long-scoreboard stalls: 6.8 of 8.4 resident warps
// in the density pass, 6.7 of 10.7 in the force pass; 42 % of the density pass's samples sit on the first use of a
// loaded candidate) while no pipe is above 50 %: a warp-wide gather is as slow as its slowest lane, and 13 % / 26 % of
// the sectors miss L1.
//
// The 32 targets of a warp are consecutive slots of the cell-sorted arrays, so for a column offset `rel` of the stencil
// the candidates of ALL of them are ONE contiguous slot range [min over lanes of run start, max over lanes of run end)
// of about 32 + 2 * reach records (instead of 32 private runs of ~9): two REDUX give its bounds, the lanes copy it with
// one or two cp.async each (16 bytes per lane, no registers, no wait), and read their own run back with LDS.128 at
// (slot - range start) one group later.  Two buffers per warp; nothing is shared between warps, so there is no block
// barrier and no warp ever waits for another one.
//
// An earlier form of this file staged per CTA (a tile of 256 targets, a producer warp issuing cp.async.bulk into an
// mbarrier ring; commit 5f398e8): its traversal loops ran almost stall-free, but the ring coupled the eight consumer
// warps (23 % of the stage waits found the data not there yet, whatever the ring depth), and the bookkeeping per group
// doubled — 0.81 / 0.81 ms against 0.73 / 0.72 ms per-lane at 1.13 M particles (profiles/r2j_staged_cta_keys.txt).
//
// A range that does not fit the buffer (targets in sparse cells next to dense ones, collapsed states) is not staged:
// the warp walks that group with the global-memory loads of pair_mask.cu.  Results never depend on which path a group
// took.
//
// Reference: SPHEngine::update_neighbor_lists' query + compute_densities + compute_pressures + compute_forces
// (src/sph_engine.cpp:335-353, 203-244).
#include "pair_stencil.cuh"

namespace sphb {

namespace {

#ifndef SPHB_STAGE_THREADS
#define SPHB_STAGE_THREADS 128
#endif
#ifndef SPHB_DSTAGE_MINBLOCKS
#define SPHB_DSTAGE_MINBLOCKS 1
#endif
#ifndef SPHB_FSTAGE_MINBLOCKS
#define SPHB_FSTAGE_MINBLOCKS 8
#endif
#ifndef SPHB_FSTAGE_CAP
#define SPHB_FSTAGE_CAP 48
#endif

#ifndef SPHB_FSTAGE_GROUP_UNROLL
#define SPHB_FSTAGE_GROUP_UNROLL 2
#endif
#ifndef SPHB_DSTAGE_GROUP_UNROLL
#define SPHB_DSTAGE_GROUP_UNROLL 2
#endif
#define SPHB_PRAGMA(x) _Pragma(#x)
#define SPHB_UNROLL_N(n) SPHB_PRAGMA(unroll n)
constexpr int kThreadsS = SPHB_STAGE_THREADS;
constexpr int kWarpsS = kThreadsS / 32;
constexpr int kColD = 64;            // density: records per staged column (<= 2 copies per lane); the record one past a
constexpr int kCapD = kColD - 1;     //          run is read (and masked) by odd tails, so ranges up to 63 are staged
constexpr int kCapF = SPHB_FSTAGE_CAP;   // force: records per staged column

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float pin(float v) { asm volatile("" : "+f"(v)); return v; }
template <typename T> __device__ __forceinline__ T* pin(T* p) { asm volatile("" : "+l"(p)); return p; }
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
// 16-byte asynchronous copy global -> shared (LDGSTS): no register staging, completion through the thread's async groups
__device__ __forceinline__ void cp_async16(const float4* dst_shared, const float4* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst_shared)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// records [lo, lo + count) of src -> dst[0 .. count), count <= 64: at most two copies per lane
__device__ __forceinline__ void stage_range(const float4* dst, const float4* __restrict__ src, uint32_t lo, uint32_t count, uint32_t lane) {
    if (lane < count) cp_async16(dst + lane, src + lo + lane);
    if (lane + 32u < count) cp_async16(dst + lane + 32u, src + lo + lane + 32u);
}

// TRUNC: the search radius cuts the kernel support short (see k_density_mask16)
template <bool SLAB, int R, bool TRUNC>
__global__ void __launch_bounds__(kThreadsS, SPHB_DSTAGE_MINBLOCKS) k_density_stage(PairArgs a) {
    __shared__ __align__(128) float4 sbuf[kWarpsS][2][2][kColD];   // [warp][buffer][column, mirror][record]
    const uint32_t lane = threadIdx.x & 31u;
    float4 (*wbuf)[2][kColD] = sbuf[threadIdx.x >> 5];
    __shared__ int4 soff[Groups<R>::kGroups + 1];                 // per group: cell offsets of {run start, run end} of column and mirror
    const size_t i_raw = (size_t)blockIdx.x * kThreadsS + threadIdx.x;
    // Lanes past the last particle walk as copies of it (no special cases inside the loops) and store nothing at the end;
    // in slab mode a warp with at least one wanted particle walks all its particles.
    const size_t i = i_raw < a.n ? i_raw : a.n - 1;
    const float4 pi = a.posm[i];
    const bool want = i_raw < a.n && (!SLAB || wants_density(a, pi));
    const GroupTable<R>& tab = group_table<R>();
    const int ng = tab.n;   // groups 0 .. ng - 1 are mirror pairs, group ng is the centre column
    {
        const int e2 = a.grid.ext[2], e12 = a.grid.ext[1] * e2;
        for (int g = threadIdx.x; g <= ng; g += kThreadsS) {
            const int rel = tab.d0[g] * e12 + tab.d1[g] * e2, reach = tab.reach[g];
            soff[g] = make_int4(rel - reach, rel + reach + 1, -rel - reach, -rel + reach + 1);
        }
    }
    __syncthreads();
    unsigned count = 0;
#ifdef SPHB_POISON_UNWANTED
    if (SLAB && !want && i_raw < a.n) {   // debug build: the records of particles whose density is not evaluated here must never be read
        const float4 v = a.velid[i];
        a.fa[i] = make_float4(pi.x, pi.y, pi.z, __int_as_float(0x7fc00000));
        a.fb[i] = make_float4(v.x, v.y, v.z, __int_as_float(0x7fc00000));
    }
#endif
    if (__any_sync(0xffffffffu, want)) {   // slab mode: whole warps of outer-halo particles leave here
        const uint32_t c = center_cell(a.grid, pi);
        const uint32_t* __restrict__ cs = pin(a.cell_start);
        const float4* __restrict__ posm = pin(a.posm);
        DensityWalker<TRUNC> dw;
        dw.init(pi, a.k, (uint32_t)(a.n >> 62));
        auto column = [&](const float4* __restrict__ src, uint32_t b, uint32_t e, uint32_t to_slot) -> uint32_t {
            return dw.column(src, b, e, to_slot, posm, pi, a.k);
        };
        // this lane's runs of group g (slots of posm)
        const uint32_t* __restrict__ csc = cs + c;
        auto bounds = [&](int g, uint32_t& b0, uint32_t& e0, uint32_t& b1, uint32_t& e1) {
            const int4 o = soff[g];
            b0 = __ldg(csc + o.x); e0 = __ldg(csc + o.y);
            b1 = __ldg(csc + o.z); e1 = __ldg(csc + o.w);
        };
        // Starts the copies of one candidate range (the union of the lanes' runs, plus one record) into `dst`; returns
        // the range start, or 0xffffffff when the range is too long to be staged.
        auto issue = [&](float4* dst, uint32_t b, uint32_t e) -> uint32_t {
            const uint32_t lo = __reduce_min_sync(0xffffffffu, b), hi = __reduce_max_sync(0xffffffffu, e);
            if (hi - lo > (uint32_t)kCapD) return 0xffffffffu;
            stage_range(dst, posm, lo, hi - lo + 1u, lane);
            return lo;
        };
        uint32_t* __restrict__ mrow = static_cast<uint32_t*>(a.masks) + i;   // (copies of the last particle rewrite its words)
        const size_t stride = a.mask_stride;

        uint32_t bA, eA, bB, eB, loA, loB;
        bounds(0, bA, eA, bB, eB);
        loA = issue(wbuf[0][0], bA, eA);
        loB = issue(wbuf[0][1], bB, eB);
        cp_async_commit();
SPHB_UNROLL_N(SPHB_DSTAGE_GROUP_UNROLL)
        for (int g = 0; g < ng; ++g) {
            uint32_t nbA, neA, nbB, neB;
            bounds(g + 1, nbA, neA, nbB, neB);   // entry ng is the centre column (both halves the same run)
            cp_async_wait_all();
            __syncwarp();   // every lane's copies of group g have landed
            float4 (*cur)[kColD] = wbuf[g & 1], (*nxt)[kColD] = wbuf[(g & 1) ^ 1];
            uint32_t word = loA != 0xffffffffu ? column(cur[0], bA - loA, eA - loA, loA) : column(posm, bA, eA, 0u);
            // the next group's ranges are copied while the mirror column is processed; their buffer was last read by
            // group g - 1 (every lane is past the __syncwarp above)
            loA = issue(nxt[0], nbA, neA);
            const uint32_t nloB = issue(nxt[1], nbB, neB);
            cp_async_commit();
            word |= (loB != 0xffffffffu ? column(cur[1], bB - loB, eB - loB, loB) : column(posm, bB, eB, 0u)) << 16;
            count += __popc(word);
            __stcs(mrow, word);
            mrow += stride;
            bA = nbA; eA = neA; bB = nbB; eB = neB; loB = nloB;
        }
        {
            cp_async_wait_all();
            __syncwarp();
            uint32_t word = loA != 0xffffffffu ? column(wbuf[ng & 1][0], bA - loA, eA - loA, loA) : column(posm, bA, eA, 0u);
            count += __popc(word);
            __stcs(mrow, word | (dw.ovf << 31));
        }
        count += dw.extra;

        if (!want) {
            count = 0;
        } else {
            const float rho = dw.density(a.k);
            const float P = a.k.gas_constant * (rho - a.k.rest_density);
            a.rho_p[i] = make_float2(rho, P);
            const float4 v = a.velid[i];
            const float A = pi.w / (2.0f * rho);
            // force-pass records, in two arrays of 16-byte halves: staged at a 16-byte stride they are gathered by
            // LDS.128 without bank conflicts (32-byte records: two-way conflicts on every access)
            a.fa[i] = make_float4(pi.x, pi.y, pi.z, A);
            a.fb[i] = make_float4(v.x, v.y, v.z, A * P);
            if (a.nbr_count) a.nbr_count[i] = count;
        }
    }
    count = __reduce_max_sync(0xffffffffu, count);
    if (lane == 0 && count > *(volatile unsigned int*)&a.sc->max_neighbors) atomicMax(&a.sc->max_neighbors, count);
}

template <bool SLAB, int R>
__global__ void __launch_bounds__(kThreadsS, SPHB_FSTAGE_MINBLOCKS) k_force_stage(PairArgs a) {
    __shared__ __align__(128) float4 sbuf[kWarpsS][2][2][2][kCapF];   // [warp][buffer][x y z A | v B][column, mirror][record]
    __shared__ int2 soff[Groups<R>::kGroups + 1];                     // per group: cell offsets of the run starts of column and mirror
    const uint32_t lane = threadIdx.x & 31u;
    float4 (*wbuf)[2][2][kCapF] = sbuf[threadIdx.x >> 5];
    const size_t i_raw = (size_t)blockIdx.x * kThreadsS + threadIdx.x;
    // Lanes past the last particle walk as copies of it (no special cases inside the loops) and store nothing at the end;
    // in slab mode the halo copies inside a warp of owned particles walk along with empty mask words.
    const size_t i = i_raw < a.n ? i_raw : a.n - 1;
    const float4 vi = a.velid[i];
    const bool want = i_raw < a.n && !(SLAB && is_ghost(vi));   // slab mode: halo copies are never advanced here
    const GroupTable<R>& tab = group_table<R>();
    const int ng = tab.n;
    const int e2 = a.grid.ext[2], e12 = a.grid.ext[1] * e2;
    for (int g = threadIdx.x; g <= ng; g += kThreadsS) {
        const int rel = tab.d0[g] * e12 + tab.d1[g] * e2, reach = tab.reach[g];
        soff[g] = make_int2(rel - reach, -rel - reach);
    }
    __syncthreads();
    if (!__any_sync(0xffffffffu, want)) return;
    const float4 pi = a.posm[i];
    const float P_i = a.rho_p[i].y;
    const uint32_t c = center_cell(a.grid, pi);
    const uint32_t* __restrict__ cs = a.cell_start;
    const uint32_t* __restrict__ csc = cs + c;
    const float4* __restrict__ fa = pin(a.fa);
    const float4* __restrict__ fb = pin(a.fb);
    const uint32_t nslots = (uint32_t)a.n;
    ForceLane fl;
    fl.init(pi, vi, P_i, a.k, (uint32_t)(a.n >> 62));
    auto eval = [&](const float4& qa, const float4& qb) { fl.eval(qa, qb); };

    const uint32_t* __restrict__ mrow = static_cast<const uint32_t*>(a.masks) + i;
    const size_t stride = a.mask_stride;
    // this lane's mask word and run starts of group g.  Candidate q of the group's column is bit 15 - q, of its mirror
    // bit 31 - q: slot = base - bit with base = run start + 15 / + 31.
    auto fetch = [&](int g, uint32_t& w, uint32_t& sA, uint32_t& sB) {
        const int2 o = soff[g];
        w = __ldcs(mrow);
        mrow += stride;
        sA = __ldg(csc + o.x);
        sB = __ldg(csc + o.y);
        if (SLAB && !want) w = 0u;
    };
    // Starts the copies of the record range one column of a group can address — [min over lanes of run start, max + 16),
    // lanes without a set bit in the column left out — into `dst` (x y z A; v B lies 2 kCapF records further);
    // returns the range start, 0xffffffff when the range is too long to be staged.
    auto issue = [&](float4* dst, bool use, uint32_t start) -> uint32_t {
        const uint32_t lo = __reduce_min_sync(0xffffffffu, use ? start : 0xffffffffu);
        const uint32_t hi = min(__reduce_max_sync(0xffffffffu, use ? start + (uint32_t)kMaskBits : 0u), nslots);
        if (hi <= lo) return 0u;             // no lane has a bit in this column
        if (hi - lo > (uint32_t)kCapF) return 0xffffffffu;
        stage_range(dst, fa, lo, hi - lo, lane);
        stage_range(dst + 2 * kCapF, fb, lo, hi - lo, lane);
        return lo;
    };

    // two groups ahead: the mask word and run starts; one group ahead: the copies
    uint32_t w, sA, sB, loA, loB, w1 = 0, sA1 = 0, sB1 = 0;
    fetch(0, w, sA, sB);
    loA = issue(&wbuf[0][0][0][0], (w & 0xFFFFu) != 0u, sA);
    loB = issue(&wbuf[0][0][1][0], (w >> 16) != 0u, sB);
    cp_async_commit();
    fetch(1, w1, sA1, sB1);   // ng >= 1 for every R
    unsigned ovf = 0;
SPHB_UNROLL_N(SPHB_FSTAGE_GROUP_UNROLL)
    for (int g = 0; g <= ng; ++g) {
        uint32_t w2 = 0, sA2 = 0, sB2 = 0, nloA = 0, nloB = 0;
        if (g + 2 <= ng) fetch(g + 2, w2, sA2, sB2);
        cp_async_wait_all();
        __syncwarp();   // every lane's copies of group g have landed; every lane is done with the other buffer
        const float4* sr = &wbuf[g & 1][0][0][0];
        if (g < ng) {
            float4* nx = &wbuf[(g & 1) ^ 1][0][0][0];
            nloA = issue(nx, (w1 & 0xFFFFu) != 0u, sA1);
            nloB = issue(nx + kCapF, g + 1 < ng && (w1 >> 16) != 0u, sB1);   // the centre word has no mirror half (bit 31: overflow flag)
            cp_async_commit();
        } else {
            ovf = w >> 31; w &= 0xFFFFu;
        }
        // every lane pops the bits of both columns as one flat stream: the warp runs max-over-lanes(popcount)
        // iterations per group, and that sum over a mirror pair is nearly lane-independent
        if (loA != 0xffffffffu && loB != 0xffffffffu) {
            const uint32_t iA = sA + 15u - loA, iB = sB + 31u - loB + (uint32_t)kCapF;
            while (w) {
                const uint32_t b = top_bit(w);
                w ^= bit_at(b);
                const float4* q = sr + ((b >= 16u ? iB : iA) - b);
                eval(q[0], q[2 * kCapF]);
            }
        } else {
            const uint32_t baseA = sA + 15u, baseB = sB + 31u;
            while (w) {
                const uint32_t b = top_bit(w);
                w ^= bit_at(b);
                const uint32_t j = (b >= 16u ? baseB : baseA) - b;
                eval(__ldg(fa + j), __ldg(fb + j));
            }
        }
        w = w1; sA = sA1; sB = sB1; loA = nloA; loB = nloB;
        w1 = w2; sA1 = sA2; sB1 = sB2;
    }
    if (!want) return;

    ForceAccum f = fl.result(a.k);
    if (ovf) {
        // some column of this particle holds more candidates than its mask has bits: the candidates beyond the mask
        // are walked with the exact radius test
        const float r2 = a.k.r2;
        auto rest = [&](uint32_t cell, uint32_t reach) {
            const uint32_t b = __ldg(cs + (cell - reach)), e = __ldg(cs + (cell + reach + 1u));
            for (uint32_t j = b + (uint32_t)kMaskBits; j < e; ++j) {
                // positions from posm: in slab mode the records are only written where the density was evaluated
                // (owned + first halo layer), which covers every ACCEPTED j of an owned particle but not every candidate
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

}  // namespace

int launch_density_stage(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    const unsigned nb = (unsigned)((a.n + kThreadsS - 1) / kThreadsS);
    const bool slab = a.slab_axis >= 0;
    // neighbor_search_radius < 2 h: the kernel support is truncated by the search radius (see k_density_mask16)
    const bool trunc = !(a.k.r2 >= 4.0f * a.k.h_sq);
#define SPHB_LAUNCH_D(RR)                                                                       \
    if (slab) { if (trunc) k_density_stage<true, RR, true><<<nb, kThreadsS, 0, st>>>(a);        \
                else k_density_stage<true, RR, false><<<nb, kThreadsS, 0, st>>>(a); }           \
    else { if (trunc) k_density_stage<false, RR, true><<<nb, kThreadsS, 0, st>>>(a);            \
           else k_density_stage<false, RR, false><<<nb, kThreadsS, 0, st>>>(a); }
    switch (a.walk_radius) {
        case 4: SPHB_LAUNCH_D(4); break;
        case 5: SPHB_LAUNCH_D(5); break;
        default: SPHB_LAUNCH_D(6); break;
    }
#undef SPHB_LAUNCH_D
    return 1;
}

int launch_force_stage(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    const unsigned nb = (unsigned)((a.n + kThreadsS - 1) / kThreadsS);
    const bool slab = a.slab_axis >= 0;
#define SPHB_LAUNCH_F(RR)                                                 \
    if (slab) k_force_stage<true, RR><<<nb, kThreadsS, 0, st>>>(a);       \
    else k_force_stage<false, RR><<<nb, kThreadsS, 0, st>>>(a)
    switch (a.walk_radius) {
        case 4: SPHB_LAUNCH_F(4); break;
        case 5: SPHB_LAUNCH_F(5); break;
        default: SPHB_LAUNCH_F(6); break;
    }
#undef SPHB_LAUNCH_F
    return 1;
}

}  // namespace sphb
