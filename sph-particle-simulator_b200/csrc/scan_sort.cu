// Device-wide single-pass scan (cell table of the per-step counting sort, neighbor.cu) and a hand-written one-sweep LSD
// radix sort (key64, val32) for sm_100a.
//
// Both serve the spatial-hash build that replaces SpatialHash::build (reference src/spatial_hash.cpp:15-25).  The
// per-step path is the counting sort of neighbor.cu, whose cell table is scanned here; the radix sort produces the
// reference-order permutation (stable sort by the reference's own (cell, id) keys) that sphb_debug_dump reports when the
// device layout differs from the reference order.
//
// HBM-bound integer work: the scan reads and writes the table once; every sort pass streams the pairs once (12 B in,
// 12 B out), staged through shared memory so the global writes of one digit are contiguous runs.
#include "sphb_internal.cuh"

namespace sphb {

namespace {

constexpr int kScanThreads = 512;
constexpr int kScanItems = kScanTile / kScanThreads;  // 8
constexpr int kSortThreads = 256;
constexpr int kSortItems = kSortTile / kSortThreads;  // 8
constexpr int kSortWarps = kSortThreads / 32;

// Lanes of the warp holding the same 8-bit digit (digit 0xFFFFFFFF = padding, matched among itself): one vote when
// the whole warp holds one value (the upper, nearly sorted digits), MATCH.ANY otherwise.  An eight-ballot
// replacement for MATCH.ANY was measured slower (sort 1.36 vs 1.26 ms at 10.7 M particles).
__device__ __forceinline__ uint32_t match_digit(uint32_t digit) {
    if (__all_sync(0xffffffffu, digit == __shfl_sync(0xffffffffu, digit, 0))) return 0xffffffffu;
    return __match_any_sync(0xffffffffu, digit);
}

template <bool MAX>
__device__ __forceinline__ uint32_t scan_op(uint32_t a, uint32_t b) {
    return MAX ? (a > b ? a : b) : a + b;
}

template <bool MAX>
__device__ __forceinline__ uint32_t warp_inclusive(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = scan_op<MAX>(v, o);
    }
    return v;
}

// Block-wide scan of one value per thread.  Returns the exclusive prefix; *total = block aggregate.
template <bool MAX, int THREADS>
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = THREADS / 32;
    uint32_t incl = warp_inclusive<MAX>(v, lane);
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < NW ? s_warp[lane] : 0u;
        uint32_t xi = warp_inclusive<MAX>(x, lane);
        if (lane < NW) s_warp[lane] = xi;
    }
    __syncthreads();
    uint32_t warp_base = w > 0 ? s_warp[w - 1] : 0u;
    uint32_t prev = __shfl_up_sync(0xffffffffu, incl, 1);
    uint32_t excl_in_warp = lane > 0 ? prev : 0u;
    *total = s_warp[NW - 1];
    return scan_op<MAX>(warp_base, excl_in_warp);
}

// ---- single-pass exclusive prefix sum with decoupled look-back ------------------------------------------------------
// Tiles take their index from an atomic ticket, so a tile only ever waits for tiles that already started.  Status word
// per tile: [63:62] flag (1 = tile aggregate, 2 = inclusive prefix), [61:0] value; one 64-bit store publishes both.
constexpr unsigned long long kScanFlagLocal = 1ull << 62, kScanFlagIncl = 2ull << 62, kScanValueMask = (1ull << 62) - 1ull;

__global__ void __launch_bounds__(kScanThreads) k_scan_excl_lookback(uint32_t* __restrict__ data, size_t n,
                                                                     volatile unsigned long long* status, unsigned int* ticket) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ unsigned int s_tile;
    __shared__ uint32_t s_prefix;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const size_t i0 = (size_t)tile * kScanTile + (size_t)tid * kScanItems;
    uint32_t v[kScanItems];
    if (i0 + kScanItems <= n) {
        const uint4 a = reinterpret_cast<const uint4*>(data + i0)[0], b = reinterpret_cast<const uint4*>(data + i0)[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) v[k] = (i0 + k < n) ? data[i0 + k] : 0u;
    }
    uint32_t run = 0, excl[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { excl[k] = run; run += v[k]; }
    uint32_t total;
    const uint32_t ex = block_exclusive<false, kScanThreads>(run, s_warp, &total);
    if (tid == 0) status[tile] = (tile == 0 ? kScanFlagIncl : kScanFlagLocal) | (unsigned long long)total;
    if (tid < 32) {
        unsigned long long prefix = 0;
        if (tile > 0) {
            long long t = (long long)tile - 1 - lane;   // lane 0 looks at the nearest predecessor
            for (;;) {
                unsigned long long w = kScanFlagIncl;   // before the first tile: an inclusive prefix of 0
                if (t >= 0) do { w = status[t]; } while ((w >> 62) == 0ull);
                const unsigned incl = __ballot_sync(0xffffffffu, (w >> 62) == 2ull);
                const int first = incl ? __ffs(incl) - 1 : 32;
                unsigned long long part = lane <= first ? (w & kScanValueMask) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                prefix += part;
                if (incl) break;
                t -= 32;
            }
            if (lane == 0) status[tile] = kScanFlagIncl | (prefix + (unsigned long long)total);
        }
        if (lane == 0) s_prefix = (uint32_t)prefix;
    }
    __syncthreads();
    const uint32_t base = s_prefix + ex;
    if (i0 + kScanItems <= n) {
        reinterpret_cast<uint4*>(data + i0)[0] = make_uint4(base + excl[0], base + excl[1], base + excl[2], base + excl[3]);
        reinterpret_cast<uint4*>(data + i0)[1] = make_uint4(base + excl[4], base + excl[5], base + excl[6], base + excl[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (i0 + k < n) data[i0 + k] = base + excl[k];
    }
}

// ---- single-pass-per-digit radix sort ("onesweep": global digit histograms up front, then one scatter kernel
// per 8-bit digit whose tiles resolve their global offsets by decoupled look-back over a status table) ----------
//
// Per pass this reads the pairs once (12 B) and writes them once (12 B) with ONE launch, instead of
// hist + 3-kernel scan + scatter (32 B, 5 launches) of the multi-kernel path above.  Tiles take their index from
// an atomic ticket, so a tile only ever waits for tiles that already started — the look-back cannot deadlock.
constexpr int kMaxPasses = 8;
constexpr unsigned long long kFlagLocal = 1ull << 62;       // status word: [63:62] flag, [61:0] count
constexpr unsigned long long kFlagIncl = 2ull << 62;
constexpr unsigned long long kCountMask = (1ull << 62) - 1ull;

struct PassPlan {
    int npasses;
    int shift[kMaxPasses];
    uint32_t mask[kMaxPasses];
};

__global__ void __launch_bounds__(kSortThreads) k_onesweep_hist(const uint64_t* __restrict__ keys, size_t n, PassPlan plan,
                                                                unsigned long long* __restrict__ ghist) {
    __shared__ uint32_t s_hist[kMaxPasses][256];
    for (int i = threadIdx.x; i < kMaxPasses * 256; i += kSortThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (size_t base = (size_t)blockIdx.x * kSortTile; base < n; base += (size_t)gridDim.x * kSortTile) {
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            const size_t e = base + (size_t)k * kSortThreads + threadIdx.x;
            const bool valid = e < n;
            const uint64_t key = valid ? keys[e] : 0ull;
            for (int p = 0; p < plan.npasses; ++p) {
                const uint32_t digit = valid ? (uint32_t)(key >> plan.shift[p]) & plan.mask[p] : 0xFFFFFFFFu;
                const uint32_t peers = match_digit(digit);
                if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_hist[p][digit], __popc(peers));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.npasses * 256; i += kSortThreads) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&ghist[i], (unsigned long long)v);
    }
}

__global__ void __launch_bounds__(kSortThreads) k_onesweep_pass(const uint64_t* __restrict__ keys_in,
                                                                const uint32_t* __restrict__ vals_in,
                                                                uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                const unsigned long long* __restrict__ ghist,
                                                                volatile unsigned long long* status, unsigned int* ticket,
                                                                size_t n, int shift, uint32_t mask) {
    __shared__ uint64_t s_keys[kSortTile];
    __shared__ uint32_t s_vals[kSortTile];
    __shared__ uint32_t s_warp_cnt[kSortWarps][256];
    __shared__ uint32_t s_tile_off[256];
    __shared__ unsigned long long s_gbase[256];
    __shared__ uint32_t s_scan[kSortWarps];
    __shared__ unsigned int s_tile;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&s_warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const unsigned int tile = s_tile;
    const size_t tile_base = (size_t)tile * kSortTile;

    uint64_t key[kSortItems];
    uint32_t val[kSortItems], lrank[kSortItems], dig[kSortItems], peers[kSortItems];
    // warp w ranks elements [w*256, w*256+256) of the tile, 32 at a time, in element order (stable).  The eight
    // MATCH.ANY of a thread are independent and issued back to back (their latency is what bounds this kernel);
    // only the running per-digit counters are accumulated round by round.
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const size_t e = tile_base + (size_t)w * (32 * kSortItems) + (size_t)k * 32 + lane;
        const bool valid = e < n;
        key[k] = valid ? keys_in[e] : ~0ull;
        val[k] = valid ? (vals_in ? vals_in[e] : (uint32_t)e) : 0u;   // first pass: the payload is the slot itself
        dig[k] = valid ? (uint32_t)(key[k] >> shift) & mask : 255u;
    }
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) peers[k] = match_digit(dig[k]);
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t below = __popc(peers[k] & ((1u << lane) - 1u));
        const int leader = __ffs(peers[k]) - 1;
        uint32_t old = 0;
        if (lane == leader) {
            old = s_warp_cnt[w][dig[k]];
            s_warp_cnt[w][dig[k]] = old + __popc(peers[k]);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        lrank[k] = old + below;
        __syncwarp();
    }
    __syncthreads();

    const size_t remaining = n - tile_base;
    const uint32_t nvalid = remaining < (size_t)kSortTile ? (uint32_t)remaining : (uint32_t)kSortTile;
    {   // thread d owns digit d
        const int d = tid;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < kSortWarps; ++ww) {
            const uint32_t t = s_warp_cnt[ww][d];
            s_warp_cnt[ww][d] = run;
            run += t;
        }
        // padding of the last tile was ranked into digit 255: it must not be published
        const uint32_t pad = (d == 255) ? (uint32_t)kSortTile - nvalid : 0u;
        const unsigned long long cnt = run - pad;
        volatile unsigned long long* my = status + (size_t)tile * 256 + d;
        if (tile == 0) *my = kFlagIncl | cnt; else *my = kFlagLocal | cnt;
        // exclusive prefix of this digit over all earlier tiles: decoupled look-back
        unsigned long long excl = 0;
        if (tile > 0) {
            long long t = (long long)tile - 1;
            for (;;) {
                unsigned long long v;
                do { v = status[(size_t)t * 256 + d]; } while ((v >> 62) == 0ull);
                excl += v & kCountMask;
                if ((v >> 62) == 2ull || t == 0) break;
                --t;
            }
            *my = kFlagIncl | (excl + cnt);
        }
        uint32_t total;
        const uint32_t ex = block_exclusive<false, kSortThreads>(run, s_scan, &total);
        s_tile_off[d] = ex;
        __syncthreads();   // s_scan is reused by the second block scan
        // global base of digit d = number of keys with a smaller digit (exclusive scan of the global histogram;
        // every prefix is <= n < 2^31) + keys of this digit in earlier tiles
        const uint32_t dbase = block_exclusive<false, kSortThreads>((uint32_t)ghist[d], s_scan, &total);
        s_gbase[d] = (unsigned long long)dbase + excl - ex;
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t pos = s_tile_off[dig[k]] + s_warp_cnt[w][dig[k]] + lrank[k];
        s_keys[pos] = key[k];
        s_vals[pos] = val[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const uint32_t pos = (uint32_t)k * kSortThreads + tid;
        if (pos < nvalid) {
            const uint64_t kk = s_keys[pos];
            const uint32_t d = (uint32_t)(kk >> shift) & mask;
            const size_t g = (size_t)(s_gbase[d] + pos);
            keys_out[g] = kk;
            vals_out[g] = s_vals[pos];
        }
    }
}

}  // namespace

// In-place exclusive prefix sum of data[0..n).  scratch holds scan_scratch_bytes(n) bytes and must be ZERO on entry
// (status words + the tile ticket); the caller clears it together with the counters it scans.
size_t scan_scratch_bytes(size_t n) { return ((n + kScanTile - 1) / kScanTile + 2) * sizeof(unsigned long long); }
int launch_scan_exclusive(uint32_t* data, size_t n, void* scratch, cudaStream_t st) {
    if (n == 0) return 0;
    const unsigned nb = (unsigned)((n + kScanTile - 1) / kScanTile);
    unsigned long long* status = static_cast<unsigned long long*>(scratch);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(status + nb);
    k_scan_excl_lookback<<<nb, kScanThreads, 0, st>>>(data, n, status, ticket);
    return 1;
}

// One-sweep variant.  scratch: (npasses * 256 + npasses * ntiles * 256) u64 + npasses u32 tickets, zeroed here.
size_t onesweep_scratch_bytes(size_t n_max, int bits_max) {
    const size_t ntiles = (n_max + kSortTile - 1) / kSortTile;
    const size_t passes = (size_t)((bits_max + 7) / 8);
    return (passes * 256 + passes * ntiles * 256) * sizeof(unsigned long long) + kMaxPasses * sizeof(unsigned int) + 64;
}

int launch_radix_sort_onesweep(const SortBuffers& sb, size_t n, int first_bit, int bits, int* out_buf, void* scratch,
                               cudaStream_t st) {
    *out_buf = 0;
    if (n == 0) return 0;
    PassPlan plan;
    plan.npasses = (bits + 7) / 8;
    if (plan.npasses > kMaxPasses) plan.npasses = kMaxPasses;
    for (int p = 0; p < plan.npasses; ++p) {
        const int nb = bits - 8 * p < 8 ? bits - 8 * p : 8;
        plan.shift[p] = first_bit + 8 * p;
        plan.mask[p] = (1u << nb) - 1u;
    }
    const uint32_t ntiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
    unsigned long long* ghist = static_cast<unsigned long long*>(scratch);
    unsigned long long* status = ghist + (size_t)plan.npasses * 256;
    unsigned int* tickets = reinterpret_cast<unsigned int*>(status + (size_t)plan.npasses * ntiles * 256);
    const size_t zero_bytes = ((size_t)plan.npasses * 256 + (size_t)plan.npasses * ntiles * 256) * sizeof(unsigned long long) +
                              kMaxPasses * sizeof(unsigned int);
    cudaMemsetAsync(scratch, 0, zero_bytes, st);
    unsigned hb = ntiles < (unsigned)(kSMs * 4) ? ntiles : (unsigned)(kSMs * 4);
    k_onesweep_hist<<<hb, kSortThreads, 0, st>>>(sb.keys[0], n, plan, ghist);
    int cur = 0, launches = 1;
    for (int p = 0; p < plan.npasses; ++p) {
        k_onesweep_pass<<<ntiles, kSortThreads, 0, st>>>(sb.keys[cur], p == 0 ? nullptr : sb.vals[cur], sb.keys[cur ^ 1], sb.vals[cur ^ 1],
                                                        ghist + (size_t)p * 256, status + (size_t)p * ntiles * 256, tickets + p, n,
                                                        plan.shift[p], plan.mask[p]);
        ++launches;
        cur ^= 1;
    }
    *out_buf = cur;
    return launches;
}

}  // namespace sphb
