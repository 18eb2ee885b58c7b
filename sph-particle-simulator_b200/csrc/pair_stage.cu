// Staged pair kernels of the default fast path (R >= 4): the neighbour-cell particles of a whole tile of targets are
// brought into shared memory by the copy engine, and the lanes traverse them from there.
//
// Same arithmetic, same neighbour masks and same summation order as the global-memory kernels of pair_mask.cu (their
// results are bit-identical; tests/test_gpu_parity.py::test_staged_equals_global); what changes is WHERE a candidate is
// read from.  In pair_mask.cu every lane fetches each of its ~600 candidates with a private LDG; ncu (profiles/
// r2b_pair_kernels.md) shows both passes waiting on those loads (long-scoreboard stalls: 6.8 of 8.4 resident warps in
// the density pass, 6.7 of 10.7 in the force pass; a warp-wide gather is as slow as its slowest lane, and 13 % / 26 %
// of the sectors miss L1, which is the tile's compulsory footprint) while no pipe is above 50 %.
//
// Here a CTA owns a TILE of kTile consecutive slots of the cell-sorted arrays.  The slots are sorted by cell, so the
// tile covers the contiguous cell range [cmin, cmax], and for a column offset `rel` of the stencil the candidates of
// ALL its targets are the particles of the cells [cmin + rel - reach, cmax + rel + reach]: ONE contiguous slot range
// [cell_start[cmin + rel - reach], cell_start[cmax + rel + reach + 1]) of about kTile + 2 * reach records, instead of
// kTile private runs of ~9.  A producer warp walks the column groups of the stencil ahead of the consumers and moves
// each group's two ranges (a column and its point mirror) into a ring of shared-memory stages with cp.async.bulk
// (1-D TMA), signalling an mbarrier per stage; the eight consumer warps (one lane per target, as before) wait on the
// stage, read their candidates with LDS.128 at (slot - range start), and release the stage through a second mbarrier.
// No __syncthreads in the loop: warps drift apart by up to kStages - 1 groups.
//
// A range that does not fit a stage (a tile that spans sparse cells next to dense ones, e.g. wall-only columns above
// the fluid surface, or collapsed states) is not staged: the header of the stage says so and the consumers walk that
// group with the global-memory loads of pair_mask.cu.  Results never depend on which path a group took.
//
// Reference: SPHEngine::update_neighbor_lists' query + compute_densities + compute_pressures + compute_forces
// (src/sph_engine.cpp:335-353, 203-244).
#include "pair_stencil.cuh"
#include "stage_pipe.cuh"

namespace sphb {

namespace {

#ifndef SPHB_STAGE_TILE
#define SPHB_STAGE_TILE 256
#endif
#ifndef SPHB_STAGE_SLACK
#define SPHB_STAGE_SLACK 64       // staged records per column beyond the tile size
#endif
#ifndef SPHB_DSTAGE_STAGES
#define SPHB_DSTAGE_STAGES 5
#endif
#ifndef SPHB_FSTAGE_STAGES
#define SPHB_FSTAGE_STAGES 2
#endif
#ifndef SPHB_DSTAGE_MINBLOCKS
#define SPHB_DSTAGE_MINBLOCKS 4
#endif
#ifndef SPHB_FSTAGE_MINBLOCKS
#define SPHB_FSTAGE_MINBLOCKS 4
#endif

constexpr int kTile = SPHB_STAGE_TILE;            // targets per CTA = consumer threads
constexpr int kConsumerWarps = kTile / 32;
constexpr int kThreadsS = kTile + 32;             // + one producer warp
constexpr int kCap = kTile + SPHB_STAGE_SLACK;    // longest candidate range that is staged
constexpr int kColD = kCap + 1;                   // density: the record one past a run is read (and masked) by odd tails
constexpr int kStagesD = SPHB_DSTAGE_STAGES;
constexpr int kStagesF = SPHB_FSTAGE_STAGES;
constexpr int kHeaderBytes = 256;                 // barriers (full[], empty[]) + stage headers
constexpr size_t kSmemD = kHeaderBytes + (size_t)kStagesD * 2 * kColD * sizeof(float4);
constexpr size_t kSmemF = kHeaderBytes + (size_t)kStagesF * 4 * kCap * sizeof(float4);
static_assert(kStagesD <= 8 && kStagesF <= 8, "header block holds at most 8 stages");

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

// Shared-memory header of a CTA: full[s] at +0, empty[s] at +64, stage headers {range start A, range start B, staged?}
// at +128.
struct StageRing {
    uint32_t bars;        // shared-space address of the header block
    uint4* hdr;
    __device__ __forceinline__ uint32_t full(int s) const { return bars + 8u * (uint32_t)s; }
    __device__ __forceinline__ uint32_t empty(int s) const { return bars + 64u + 8u * (uint32_t)s; }
};

// The producer warp.  Lane l looks up the range bounds of groups l, l + 32, ... once (one round trip to the cell
// table for the whole tile), then lane 0 feeds the ring: wait until the consumers have released the stage, publish
// the range starts, and either issue the bulk copies (arming the stage's barrier with their byte count) or mark the
// group as not staged.  NSRC arrays of 16-byte records are staged (density: posm; force: the two record halves), a
// stage being laid out as [src][column A | column B][COL_RECS]; EXTRA: records copied past each range.
template <int R, int NS, int NSRC, int EXTRA, int COL_RECS>
__device__ __forceinline__ void produce(const PairArgs& a, const StageRing& ring, uint32_t stage0, const float4* src0,
                                        const float4* src1, size_t i0) {
    const int lane = threadIdx.x & 31;
    const GroupTable<R>& tab = group_table<R>();
    const int ng = tab.n;   // groups 0 .. ng - 1 are mirror pairs, group ng is the centre column
    const int e2 = a.grid.ext[2], e12 = a.grid.ext[1] * e2;
    const uint32_t nt = (uint32_t)min((size_t)kTile, a.n - i0);
    const uint32_t cmin = center_cell(a.grid, a.posm[i0]), cmax = center_cell(a.grid, a.posm[i0 + nt - 1]);
    constexpr int kSlots = (Groups<R>::kGroups + 1 + 31) / 32;
    uint32_t A0[kSlots], A1[kSlots], B0[kSlots], B1[kSlots];
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {
        const int g = lane + 32 * k;
        A0[k] = A1[k] = B0[k] = B1[k] = 0u;
        if (g <= ng) {
            const uint32_t rel = (uint32_t)(tab.d0[g] * e12 + tab.d1[g] * e2), reach = (uint32_t)tab.reach[g];
            A0[k] = __ldg(a.cell_start + (cmin + rel - reach)); A1[k] = __ldg(a.cell_start + (cmax + rel + reach + 1u));
            B0[k] = __ldg(a.cell_start + (cmin - rel - reach)); B1[k] = __ldg(a.cell_start + (cmax - rel + reach + 1u));
        }
    }
    int s = 0;
    uint32_t parity = 1u;   // a fresh barrier counts as having completed the phase of parity 1: the first round does not wait
    for (int g = 0; g <= ng; ++g) {
        uint32_t va0 = A0[0], va1 = A1[0], vb0 = B0[0], vb1 = B1[0];
#pragma unroll
        for (int k = 1; k < kSlots; ++k)
            if ((g >> 5) == k) { va0 = A0[k]; va1 = A1[k]; vb0 = B0[k]; vb1 = B1[k]; }
        const uint32_t a0 = __shfl_sync(0xffffffffu, va0, g & 31), a1 = __shfl_sync(0xffffffffu, va1, g & 31);
        const uint32_t b0 = __shfl_sync(0xffffffffu, vb0, g & 31), b1 = __shfl_sync(0xffffffffu, vb1, g & 31);
        if (lane == 0) {
            mbar_wait(ring.empty(s), parity);
            const uint32_t lenA = a1 - a0, lenB = (g == ng) ? 0u : b1 - b0;
            const bool fits = lenA <= (uint32_t)kCap && lenB <= (uint32_t)kCap;
            ring.hdr[s] = make_uint4(a0, b0, fits ? 1u : 0u, 0u);
            const uint32_t bytesA = fits && (lenA + EXTRA) ? (lenA + EXTRA) * 16u : 0u;
            const uint32_t bytesB = fits && g != ng && (lenB + EXTRA) ? (lenB + EXTRA) * 16u : 0u;
            if (bytesA + bytesB) {
                mbar_arrive_expect_tx(ring.full(s), NSRC * (bytesA + bytesB));
                const uint32_t dst = stage0 + (uint32_t)s * (NSRC * 2u * COL_RECS * 16u);
                if (bytesA) bulk_g2s(dst, src0 + a0, bytesA, ring.full(s));
                if (bytesB) bulk_g2s(dst + COL_RECS * 16u, src0 + b0, bytesB, ring.full(s));
                if (NSRC == 2) {
                    if (bytesA) bulk_g2s(dst + 2u * COL_RECS * 16u, src1 + a0, bytesA, ring.full(s));
                    if (bytesB) bulk_g2s(dst + 3u * COL_RECS * 16u, src1 + b0, bytesB, ring.full(s));
                }
            } else {
                mbar_arrive(ring.full(s));
            }
        }
        if (++s == NS) { s = 0; parity ^= 1u; }
    }
}

__device__ __forceinline__ StageRing ring_init(unsigned char* smem, int stages) {
    StageRing ring;
    ring.bars = smem_addr(smem);
    ring.hdr = reinterpret_cast<uint4*>(smem + 128);
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(ring.full(s), 1u); mbar_init(ring.empty(s), (uint32_t)kConsumerWarps); }
        mbar_fence_init();
    }
    return ring;
}

// TRUNC: the search radius cuts the kernel support short (see k_density_mask16)
template <bool SLAB, int R, bool TRUNC>
__global__ void __launch_bounds__(kThreadsS, SPHB_DSTAGE_MINBLOCKS) k_density_stage(PairArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const StageRing ring = ring_init(smem, kStagesD);
    const float4* stage0 = reinterpret_cast<const float4*>(smem + kHeaderBytes);
    const int lane = threadIdx.x & 31;
    const bool consumer = threadIdx.x < kTile;
    const size_t i0 = (size_t)blockIdx.x * kTile;
    const size_t i = i0 + threadIdx.x;
    float4 pi = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    bool want = false;
    if (consumer && i < a.n) {
        pi = a.posm[i];
        want = !SLAB || wants_density(a, pi);
    }
    if (!__syncthreads_or(want ? 1 : 0)) return;   // slab mode: a tile of outer-halo particles only; also publishes the barriers
    if (!consumer) {
        produce<R, kStagesD, 1, 1, kColD>(a, ring, smem_addr(stage0), a.posm, nullptr, i0);
        return;
    }

    unsigned count = 0;
    const uint32_t c = want ? center_cell(a.grid, pi) : 0u;
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

    // squared distances of two candidates to this particle with the reference's roundings (spatial_hash.h:70-73):
    // fl(fl(fl(dx dx) + fl(dy dy)) + fl(dz dz)); squares as fma(d, d, -0) (see pair_mask.cu)
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
    uint32_t m;
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
    };
    // One cell column: candidates [b, e) of `src` (shared-memory stage or posm itself); `to_slot` turns an index of
    // src into a slot of posm.  Returns the 16-bit mask, candidate q at bit 15 - q.
    auto column = [&](const float4* __restrict__ src, uint32_t b, uint32_t e, uint32_t to_slot) -> uint32_t {
        const uint32_t end = min(e, b + (uint32_t)kMaskBits);
        m = 0;
        uint32_t j = b;
#pragma unroll 1
        for (; j < end; j += 2) {
            // index `end` may be read (as the masked second half of an odd tail): the stages hold one record past
            // their range and slot n of posm is a finite sentinel
            const float4 pa = src[j], pb = src[j + 1];
            visit(pa, pb, j + 1 >= end);
        }
        m <<= (b + (uint32_t)kMaskBits) - j;   // j - b candidates were shifted in (an even number <= 16)
        if (e > end) {   // more than 16 candidates in this column: no mask for the rest
            ovf = 1u;
            const float r2 = a.k.r2, inv_h = a.k.inv_h;
            for (uint32_t u = end + to_slot; u < e + to_slot; ++u) {
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

    const GroupTable<R>& tab = group_table<R>();
    const int ng = tab.n;
    const int e2 = a.grid.ext[2], e12 = a.grid.ext[1] * e2;
    uint32_t bA = 0, eA = 0, bB = 0, eB = 0;
    // this lane's runs of group g (slots of posm); issued one group ahead of their use
    auto bounds = [&](int g, uint32_t& b0, uint32_t& e0, uint32_t& b1, uint32_t& e1) {
        const uint32_t rel = (uint32_t)(tab.d0[g] * e12 + tab.d1[g] * e2), reach = (uint32_t)tab.reach[g];
        b0 = __ldg(cs + (c + rel - reach)); e0 = __ldg(cs + (c + rel + reach + 1u));
        b1 = __ldg(cs + (c - rel - reach)); e1 = __ldg(cs + (c - rel + reach + 1u));
    };
    if (want) bounds(0, bA, eA, bB, eB);
    uint32_t* __restrict__ mrow = static_cast<uint32_t*>(a.masks) + i;
    const size_t stride = a.mask_stride;
    int s = 0;
    uint32_t parity = 0u;
#pragma unroll 1
    for (int g = 0; g <= ng; ++g) {
        uint32_t nbA = 0, neA = 0, nbB = 0, neB = 0;
        if (want && g < ng) bounds(g + 1, nbA, neA, nbB, neB);   // entry ng is the centre column (both halves the same run)
        mbar_wait(ring.full(s), parity);
        const uint4 h = ring.hdr[s];
        if (want) {
            uint32_t word;
            if (h.z) {
                const float4* sA = stage0 + (size_t)s * (2 * kColD);
                word = column(sA, bA - h.x, eA - h.x, h.x);
                if (g < ng) word |= column(sA + kColD, bB - h.y, eB - h.y, h.y) << 16;
            } else {
                word = column(posm, bA, eA, 0u);
                if (g < ng) word |= column(posm, bB, eB, 0u) << 16;
            }
            count += __popc(word);
            if (g == ng) word |= ovf << 31;
            __stcs(mrow, word);
            mrow += stride;
        }
        bA = nbA; eA = neA; bB = nbB; eB = neB;
        __syncwarp();
        if (lane == 0) mbar_arrive(ring.empty(s));
        if (++s == kStagesD) { s = 0; parity ^= 1u; }
    }

    if (want) {
        const float rho = (rho0 + rho1) * (a.k.sigma * (1.0f / 6.0f));
        const float P = a.k.gas_constant * (rho - a.k.rest_density);
        a.rho_p[i] = make_float2(rho, P);
        const float4 v = a.velid[i];
        const float A = pi.w / (2.0f * rho);
        // force-pass records, in two arrays of 16-byte halves: a stage then holds them at a 16-byte stride, which the
        // lanes' LDS.128 gathers read without bank conflicts (32-byte records: two-way conflicts on every access)
        a.fa[i] = make_float4(pi.x, pi.y, pi.z, A);
        a.fb[i] = make_float4(v.x, v.y, v.z, A * P);
        if (a.nbr_count) a.nbr_count[i] = count;
    }
    count = __reduce_max_sync(0xffffffffu, count);
    if (lane == 0 && count > *(volatile unsigned int*)&a.sc->max_neighbors) atomicMax(&a.sc->max_neighbors, count);
}

template <bool SLAB, int R>
__global__ void __launch_bounds__(kThreadsS, SPHB_FSTAGE_MINBLOCKS) k_force_stage(PairArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const StageRing ring = ring_init(smem, kStagesF);
    const float4* stage0 = reinterpret_cast<const float4*>(smem + kHeaderBytes);
    const int lane = threadIdx.x & 31;
    const bool consumer = threadIdx.x < kTile;
    const size_t i0 = (size_t)blockIdx.x * kTile;
    const size_t i = i0 + threadIdx.x;
    float4 vi = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    bool want = false;
    if (consumer && i < a.n) {
        vi = a.velid[i];
        want = !(SLAB && is_ghost(vi));   // slab mode: halo copies are never advanced here
    }
    if (!__syncthreads_or(want ? 1 : 0)) return;
    if (!consumer) {
        produce<R, kStagesF, 2, 0, kCap>(a, ring, smem_addr(stage0), a.fa, a.fb, i0);
        return;
    }

    const float4 pi = want ? a.posm[i] : make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    const float P_i = want ? a.rho_p[i].y : 0.0f;
    const uint32_t c = want ? center_cell(a.grid, pi) : 0u;
    const uint32_t* __restrict__ cs = a.cell_start;
    const float4* __restrict__ fa = pin(a.fa);
    const float4* __restrict__ fb = pin(a.fb);
    const float2 npxy = f2(-pi.x, -pi.y), nvxy = f2(-vi.x, -vi.y);
    const float ninvh = pin(-a.k.inv_h);
    // accumulators of -F_pressure / (sigma / h) and F_viscosity / (2 mu sigma / h^2): (x, y) packed, z scalar
    float2 fpxy = f2(0.0f, 0.0f), fvxy = f2(0.0f, 0.0f);
    float fpz = 0.0f, fvz = 0.0f;
    // pair j -> i without a distance test (j was accepted by the density pass); see k_force_mask16 for the derivation
    auto eval = [&](const float4& qa, const float4& qb) {   // qa = {x, y, z, A}, qb = {vx, vy, vz, B}
        const float2 rxy = __fadd2_rn(f2(qa.x, qa.y), npxy);
        const float rz = qa.z - pi.z;
        const float d2 = fmaf(rz, rz, fmaf(rxy.y, rxy.y, rxy.x * rxy.x));
        const float inv_len = fast_rsqrt(fmaxf(d2, 1e-30f));
        const float t2 = fmaxf(fmaf(d2 * ninvh, inv_len, 2.0f), 0.0f);   // (2 - q)+
        const float t1 = fmaxf(t2 - 1.0f, 0.0f);                          // (1 - q)+
        const float gh = fmaf(-0.25f * t2, t2, t1 * t1);                  // dW/dq / (2 sigma) = (1-q)+^2 - (2-q)+^2 / 4
        const float lq = fmaf(-4.0f, t1, t2);                             // d2W/dq2 / sigma
        const float cp = fmaf(qa.w, P_i, qb.w) * (gh * inv_len);
        fpxy = __ffma2_rn(f2(cp, cp), rxy, fpxy);
        fpz = fmaf(cp, rz, fpz);
        const float cv = qa.w * lq;
        const float2 uxy = __fadd2_rn(f2(qb.x, qb.y), nvxy);
        fvxy = __ffma2_rn(f2(cv, cv), uxy, fvxy);
        fvz = fmaf(cv, qb.z - vi.z, fvz);
    };

    const GroupTable<R>& tab = group_table<R>();
    const int ng = tab.n;
    const int e2 = a.grid.ext[2], e12 = a.grid.ext[1] * e2;
    const uint32_t* __restrict__ mrow = static_cast<const uint32_t*>(a.masks) + i;
    const size_t stride = a.mask_stride;
    // this lane's mask word and run starts of group g; issued one group ahead of their use.  Candidate q of the
    // group's column is bit 15 - q, of its mirror bit 31 - q: slot = base - bit.
    auto fetch = [&](int g, uint32_t& w, uint32_t& baseA, uint32_t& baseB) {
        const uint32_t rel = (uint32_t)(tab.d0[g] * e12 + tab.d1[g] * e2), reach = (uint32_t)tab.reach[g];
        w = __ldcs(mrow);
        mrow += stride;
        baseA = __ldg(cs + (c + rel - reach)) + 15u;
        baseB = __ldg(cs + (c - rel - reach)) + 31u;
    };
    uint32_t w = 0, baseA = 0, baseB = 0;
    if (want) fetch(0, w, baseA, baseB);
    unsigned ovf = 0;
    int s = 0;
    uint32_t parity = 0u;
#pragma unroll 1
    for (int g = 0; g <= ng; ++g) {
        uint32_t nw = 0, nbaseA = 0, nbaseB = 0;
        if (want && g < ng) fetch(g + 1, nw, nbaseA, nbaseB);
        if (g == ng) { ovf = w >> 31; w &= 0xFFFFu; }
        mbar_wait(ring.full(s), parity);
        const uint4 h = ring.hdr[s];
        // every lane pops the bits of both columns as one flat stream: the warp runs max-over-lanes(popcount)
        // iterations per group, and that sum over a mirror pair is nearly lane-independent
        if (h.z) {
            const float4* sr = stage0 + (size_t)s * (4 * kCap);   // [x y z A: column, mirror][v B: column, mirror]
            const uint32_t iA = baseA - h.x, iB = baseB - h.y + (uint32_t)kCap;
            while (w) {
                const uint32_t b = top_bit(w);
                w ^= bit_at(b);
                const float4* q = sr + ((b >= 16u ? iB : iA) - b);
                eval(q[0], q[2 * kCap]);
            }
        } else {
            while (w) {
                const uint32_t b = top_bit(w);
                w ^= bit_at(b);
                const uint32_t j = (b >= 16u ? baseB : baseA) - b;
                eval(__ldg(fa + j), __ldg(fb + j));
            }
        }
        w = nw; baseA = nbaseA; baseB = nbaseB;
        __syncwarp();
        if (lane == 0) mbar_arrive(ring.empty(s));
        if (++s == kStagesF) { s = 0; parity ^= 1u; }
    }
    if (!want) return;

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
        auto rest = [&](uint32_t cell, uint32_t reach) {
            const uint32_t b = __ldg(cs + (cell - reach)), e = __ldg(cs + (cell + reach + 1u));
            for (uint32_t j = b + (uint32_t)kMaskBits; j < e; ++j) {
                // positions from posm: in slab mode the records are only written where the density was evaluated (owned +
                // first halo layer), which covers every ACCEPTED j of an owned particle but not every candidate
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

// Function attributes are per device: `done` remembers the devices this kernel has been configured on.
template <typename K>
int prepare(K kernel, size_t smem_bytes, unsigned long long* done) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (*done >> dev & 1ull) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return -1;
    *done |= 1ull << dev;
    return 0;
}

}  // namespace

// Both return the number of kernels enqueued, or -1 when the staged kernels cannot be configured on this device.
int launch_density_stage(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    const unsigned nb = (unsigned)((a.n + kTile - 1) / kTile);
    const bool slab = a.slab_axis >= 0;
    const bool trunc = !(a.k.r2 >= 4.0f * a.k.h_sq);
#define SPHB_GO(K) do { static unsigned long long done = 0; if (prepare(K, kSmemD, &done)) return -1; K<<<nb, kThreadsS, kSmemD, st>>>(a); } while (0)
#define SPHB_LAUNCH_D(RR)                                                                                      \
    if (slab) { if (trunc) SPHB_GO((k_density_stage<true, RR, true>)); else SPHB_GO((k_density_stage<true, RR, false>)); }    \
    else { if (trunc) SPHB_GO((k_density_stage<false, RR, true>)); else SPHB_GO((k_density_stage<false, RR, false>)); }
    switch (a.walk_radius) {
        case 4: SPHB_LAUNCH_D(4); break;
        case 5: SPHB_LAUNCH_D(5); break;
        default: SPHB_LAUNCH_D(6); break;
    }
#undef SPHB_LAUNCH_D
#undef SPHB_GO
    return 1;
}

int launch_force_stage(const PairArgs& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    const unsigned nb = (unsigned)((a.n + kTile - 1) / kTile);
    const bool slab = a.slab_axis >= 0;
#define SPHB_GO(K) do { static unsigned long long done = 0; if (prepare(K, kSmemF, &done)) return -1; K<<<nb, kThreadsS, kSmemF, st>>>(a); } while (0)
#define SPHB_LAUNCH_F(RR) if (slab) SPHB_GO((k_force_stage<true, RR>)); else SPHB_GO((k_force_stage<false, RR>))
    switch (a.walk_radius) {
        case 4: SPHB_LAUNCH_F(4); break;
        case 5: SPHB_LAUNCH_F(5); break;
        default: SPHB_LAUNCH_F(6); break;
    }
#undef SPHB_LAUNCH_F
#undef SPHB_GO
    return 1;
}

}  // namespace sphb
