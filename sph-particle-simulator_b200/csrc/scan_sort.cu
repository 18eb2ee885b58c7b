// Device-wide scan and hand-written LSD radix sort (key64, val32) for sm_100a.
//
// Replaces the serial unordered_map build of SpatialHash::build (reference src/spatial_hash.cpp:15-25):
// instead of pushing particle ids into per-cell vectors, particles are stably sorted by
// (cell, id) so every cell's members are contiguous and in ascending id — the order in which
// the reference's per-cell vectors hold them.
//
// HBM-bound integer work: every pass streams the pairs once for the digit histogram (8 B/pair) and
// once for the scatter (12 B in, 12 B out), staged through shared memory so the global writes of
// one digit are contiguous runs.
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

template <bool MAX>
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const uint32_t* __restrict__ in, size_t n,
                                                              uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const size_t base = (size_t)blockIdx.x * kScanTile;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        size_t i = base + (size_t)k * kScanThreads + threadIdx.x;
        if (i < n) acc = scan_op<MAX>(acc, in[i]);
    }
    uint32_t total;
    block_exclusive<MAX, kScanThreads>(acc, s_warp, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// Single CTA: exclusive scan of the per-block aggregates, in place.
template <bool MAX>
__global__ void __launch_bounds__(1024) k_scan_block_sums(uint32_t* __restrict__ bs, size_t nb) {
    __shared__ uint32_t s_warp[32];
    uint32_t carry = 0;
    for (size_t base = 0; base < nb; base += 1024) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < nb ? bs[i] : 0u;
        uint32_t total;
        uint32_t ex = block_exclusive<MAX, 1024>(v, s_warp, &total);
        if (i < nb) bs[i] = scan_op<MAX>(carry, ex);
        carry = scan_op<MAX>(carry, total);
        __syncthreads();
    }
}

template <bool MAX>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const uint32_t* in, uint32_t* out, size_t n,
                                                             const uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const size_t i0 = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
    uint32_t run = 0;
    uint32_t incl[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        run = scan_op<MAX>(run, v[k]);
        incl[k] = run;
    }
    uint32_t total;
    uint32_t ex = block_exclusive<MAX, kScanThreads>(run, s_warp, &total);
    const uint32_t base = scan_op<MAX>(block_sums[blockIdx.x], ex);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (i0 + k < n) {
            if (MAX) out[i0 + k] = scan_op<true>(base, incl[k]);
            else out[i0 + k] = base + (incl[k] - v[k]);
        }
    }
}

template <bool MAX>
int launch_scan(uint32_t* data, size_t n, uint32_t* block_sums, cudaStream_t st) {
    if (n == 0) return 0;
    const unsigned nb = (unsigned)((n + kScanTile - 1) / kScanTile);
    k_scan_reduce<MAX><<<nb, kScanThreads, 0, st>>>(data, n, block_sums);
    k_scan_block_sums<MAX><<<1, 1024, 0, st>>>(block_sums, nb);
    k_scan_apply<MAX><<<nb, kScanThreads, 0, st>>>(data, data, n, block_sums);
    return 3;
}

// ---- radix sort ---------------------------------------------------------------------------------

__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const uint64_t* __restrict__ keys, size_t n, int shift,
                                                             uint32_t mask, uint32_t ntiles,
                                                             uint32_t* __restrict__ counts) {
    __shared__ uint32_t s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * kSortTile;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        size_t e = base + (size_t)k * kSortThreads + threadIdx.x;
        bool valid = e < n;
        uint32_t digit = valid ? (uint32_t)(keys[e] >> shift) & mask : 0xFFFFFFFFu;
        // keys arrive nearly sorted, so most lanes of a warp share a digit: aggregate before the atomic
        uint32_t peers = __match_any_sync(0xffffffffu, digit);
        if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_hist[digit], __popc(peers));
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * ntiles + blockIdx.x] = s_hist[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads) k_radix_scatter(const uint64_t* __restrict__ keys_in,
                                                                const uint32_t* __restrict__ vals_in,
                                                                uint64_t* __restrict__ keys_out,
                                                                uint32_t* __restrict__ vals_out,
                                                                const uint32_t* __restrict__ scanned, size_t n, int shift,
                                                                uint32_t mask, uint32_t ntiles) {
    __shared__ uint64_t s_keys[kSortTile];
    __shared__ uint32_t s_vals[kSortTile];
    __shared__ uint32_t s_warp_cnt[kSortWarps][256];
    __shared__ uint32_t s_tile_off[256];
    __shared__ uint32_t s_gbase[256];
    __shared__ uint32_t s_scan[kSortWarps];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const size_t tile_base = (size_t)blockIdx.x * kSortTile;

    for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&s_warp_cnt[0][0])[i] = 0;
    __syncthreads();

    uint64_t key[kSortItems];
    uint32_t val[kSortItems];
    uint32_t lrank[kSortItems];
    uint32_t dig[kSortItems];
    // warp w ranks elements [w*256, w*256+256) of the tile, 32 at a time, in element order (stable)
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        size_t e = tile_base + (size_t)w * (32 * kSortItems) + (size_t)k * 32 + lane;
        bool valid = e < n;
        key[k] = valid ? keys_in[e] : ~0ull;
        val[k] = valid ? vals_in[e] : 0u;
        uint32_t digit = valid ? (uint32_t)(key[k] >> shift) & mask : 255u;  // padding ranks last in its bin
        dig[k] = digit;
        uint32_t peers = __match_any_sync(0xffffffffu, digit);
        uint32_t below = __popc(peers & ((1u << lane) - 1u));
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) {
            old = s_warp_cnt[w][digit];
            s_warp_cnt[w][digit] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        lrank[k] = old + below;
        __syncwarp();
    }
    __syncthreads();

    {   // thread d owns digit d: prefix over warps, then exclusive scan over digits
        const int d = tid;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < kSortWarps; ++ww) {
            uint32_t t = s_warp_cnt[ww][d];
            s_warp_cnt[ww][d] = run;
            run += t;
        }
        uint32_t total;
        uint32_t ex = block_exclusive<false, kSortThreads>(run, s_scan, &total);
        s_tile_off[d] = ex;
        s_gbase[d] = scanned[(size_t)d * ntiles + blockIdx.x] - ex;
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        uint32_t pos = s_tile_off[dig[k]] + s_warp_cnt[w][dig[k]] + lrank[k];
        s_keys[pos] = key[k];
        s_vals[pos] = val[k];
    }
    __syncthreads();

    const size_t remaining = n - tile_base;
    const uint32_t nvalid = remaining < (size_t)kSortTile ? (uint32_t)remaining : (uint32_t)kSortTile;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        uint32_t pos = (uint32_t)k * kSortThreads + tid;
        if (pos < nvalid) {
            uint64_t kk = s_keys[pos];
            uint32_t d = (uint32_t)(kk >> shift) & mask;
            uint32_t g = s_gbase[d] + pos;
            keys_out[g] = kk;
            vals_out[g] = s_vals[pos];
        }
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

int launch_scan_sum_exclusive(uint32_t* data, size_t n, uint32_t* block_sums, cudaStream_t st) {
    return launch_scan<false>(data, n, block_sums, st);
}
int launch_scan_max_inclusive(uint32_t* data, size_t n, uint32_t* block_sums, cudaStream_t st) {
    return launch_scan<true>(data, n, block_sums, st);
}

int launch_radix_sort(const SortBuffers& sb, size_t n, int bits, int* out_buf, cudaStream_t st) {
    int cur = 0, launches = 0;
    if (n == 0) { *out_buf = 0; return 0; }
    const uint32_t ntiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
    for (int shift = 0; shift < bits; shift += 8) {
        const int nb = bits - shift < 8 ? bits - shift : 8;
        const uint32_t mask = (1u << nb) - 1u;
        k_radix_hist<<<ntiles, kSortThreads, 0, st>>>(sb.keys[cur], n, shift, mask, ntiles, sb.counts);
        launches += 1;
        launches += launch_scan<false>(sb.counts, (size_t)256 * ntiles, sb.block_sums, st);
        k_radix_scatter<<<ntiles, kSortThreads, 0, st>>>(sb.keys[cur], sb.vals[cur], sb.keys[cur ^ 1], sb.vals[cur ^ 1],
                                                        sb.counts, n, shift, mask, ntiles);
        launches += 1;
        cur ^= 1;
    }
    *out_buf = cur;
    return launches;
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
