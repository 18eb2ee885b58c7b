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

}  // namespace sphb
