// Slab decomposition support: migration / halo packing, record append, owned-particle export.
//
// No reference counterpart — the reference is a single-process CPU program.  One
// context per GPU owns the reference cells [own_lo, own_hi) along one axis; copies of the neighbours'
// boundary layers ("ghosts", id word bit 31 set) are appended before a step so that every owned
// particle sees its complete neighbourhood.  Packing order is irrelevant by construction: after an
// exchange the step re-sorts owned + ghost particles by (cell, global id), so the device layout — and
// with it every summation order — is independent of how many GPUs share the domain.
//
// Exchange record: 32 bytes = {x, y, z, mass} {vx, vy, vz, id}; exactly one posm + one velid entry.
#include "sphb_internal.cuh"

namespace sphb {

namespace {

constexpr int kThreads = 256;
constexpr unsigned kGhostBit = 0x80000000u;
constexpr unsigned kFull = 0xffffffffu;

inline unsigned blocks_for(size_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

// position on the slab axis in units of reference cells; its floor is the reference cell (spatial_hash.h:30-36)
__device__ __forceinline__ float axis_scaled(const float4& p, int axis, float ref_inv_cell) {
    const float c = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
    return __fmul_rn(c, ref_inv_cell);
}

// destination rank of a reference cell: the interval [cuts[d], cuts[d+1]) containing it (clamped to the ends)
__device__ __forceinline__ int dest_of(const SlabCuts& sc, int cell) {
    int lo = 0, hi = sc.nranks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (cell >= sc.cuts[mid]) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// warp-aggregated append: returns this lane's slot in the group `key` (all lanes call; inactive lanes pass key = -1)
__device__ __forceinline__ unsigned grouped_slot(int key, unsigned int* counters) {
    const unsigned peers = __match_any_sync(kFull, key);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (key >= 0 && lane == leader) base = atomicAdd(&counters[key], __popc(peers));
    base = __shfl_sync(kFull, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(kThreads) k_slab_append(size_t count, const float4* __restrict__ rec, unsigned flag,
                                                          float4* __restrict__ posm, float4* __restrict__ velid) {
    const size_t k = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (k >= count) return;
    posm[k] = rec[2 * k];
    float4 v = rec[2 * k + 1];
    v.w = __uint_as_float((__float_as_uint(v.w) & ~kGhostBit) | flag);
    velid[k] = v;
}

__global__ void __launch_bounds__(kThreads) k_slab_export(size_t n, const float4* __restrict__ posm,
                                                          const float4* __restrict__ velid, const float2* __restrict__ rho_p,
                                                          const float4* __restrict__ acc, uint32_t* __restrict__ ids,
                                                          float* __restrict__ pos3, float* __restrict__ vel3,
                                                          float* __restrict__ rho, float* __restrict__ P,
                                                          float* __restrict__ acc3, unsigned int* __restrict__ cursor) {
    const size_t s = (size_t)blockIdx.x * kThreads + threadIdx.x;
    int key = -1;
    float4 v;
    if (s < n) {
        v = velid[s];
#ifdef SPHB_DEBUG_EXPORT_GHOSTS   // debug builds only (tools/debug): halo copies are exported too, their ids keep bit 31
        key = 0;
#else
        if (!(__float_as_uint(v.w) & kGhostBit)) key = 0;
#endif
    }
    const size_t k = grouped_slot(key, cursor);
    if (key < 0) return;
    ids[k] = __float_as_uint(v.w);
    if (pos3) { const float4 p = posm[s]; pos3[3 * k] = p.x; pos3[3 * k + 1] = p.y; pos3[3 * k + 2] = p.z; }
    if (vel3) { vel3[3 * k] = v.x; vel3[3 * k + 1] = v.y; vel3[3 * k + 2] = v.z; }
    if (rho || P) { const float2 rp = rho_p[s]; if (rho) rho[k] = rp.x; if (P) P[k] = rp.y; }
    if (acc3) { const float4 a = acc[s]; acc3[3 * k] = a.x; acc3[3 * k + 1] = a.y; acc3[3 * k + 2] = a.z; }
}

__global__ void __launch_bounds__(kThreads) k_pack_upload_ids(size_t n, const float* __restrict__ pos3,
                                                              const float* __restrict__ vel3, const float* __restrict__ mass,
                                                              const uint32_t* __restrict__ ids, float default_mass,
                                                              float4* __restrict__ posm, float4* __restrict__ velid) {
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    posm[i] = make_float4(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2], mass ? mass[i] : default_mass);
    velid[i] = make_float4(vel3 ? vel3[3 * i] : 0.0f, vel3 ? vel3[3 * i + 1] : 0.0f, vel3 ? vel3[3 * i + 2] : 0.0f,
                           __uint_as_float(ids[i] & ~kGhostBit));
}

// ---- one-round exchange: every owned particle goes to its (possibly new) owner and, flagged as ghost,
// to each adjacent rank whose halo layers contain its cell.  Keys: 2*r = "owned by r", 2*r+1 = "ghost for r".
struct Route { int owner, ghost_lo, ghost_hi; };   // ranks; -1 = none

// A neighbour needs the particles of `layers` reference cells beyond its face — and a sliver more: the refined internal
// cells are 0.1 % larger than neighbor_search_radius / refine (make_grid), so the stencil walk of a first-layer halo
// particle reaches up to 0.002 cells past the last layer.  Candidates there are always rejected by the radius test, but
// the packed density pass splits its sum over even / odd candidates of a run, so a rejected candidate at the start of a
// run still decides which partial sum its successors enter: without the sliver a halo particle's density could differ
// from the single-context one in the last bit once particles drift off the lattice planes.  kHaloSliver = 1/64 cell.
#ifndef SPHB_HALO_SLIVER
#define SPHB_HALO_SLIVER 0.015625f
#endif
constexpr float kHaloSliver = SPHB_HALO_SLIVER;

__device__ __forceinline__ Route route_of(const SlabCuts& sc, float scaled, int layers) {
    Route r;
    const int cell = __float2int_rd(scaled);
    r.owner = dest_of(sc, cell);
    r.ghost_lo = (r.owner > 0 && scaled < (float)(sc.cuts[r.owner] + layers) + kHaloSliver) ? r.owner - 1 : -1;
    r.ghost_hi = (r.owner < sc.nranks - 1 && scaled >= (float)(sc.cuts[r.owner + 1] - layers) - kHaloSliver) ? r.owner + 1 : -1;
    return r;
}

__device__ __forceinline__ void grouped_count(int key, unsigned int* counters) {
    const unsigned peers = __match_any_sync(kFull, key);
    if (key >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&counters[key], __popc(peers));
}

__global__ void __launch_bounds__(kThreads) k_exchange_count(size_t n, const float4* __restrict__ posm,
                                                             const float4* __restrict__ velid, SlabCuts sc, int axis,
                                                             float ref_inv_cell, int layers, unsigned int* __restrict__ counts) {
    const size_t s = (size_t)blockIdx.x * kThreads + threadIdx.x;
    Route r = {-1, -1, -1};
    if (s < n && !(__float_as_uint(velid[s].w) & kGhostBit)) r = route_of(sc, axis_scaled(posm[s], axis, ref_inv_cell), layers);
    grouped_count(r.owner >= 0 ? 2 * r.owner : -1, counts);
    grouped_count(r.ghost_lo >= 0 ? 2 * r.ghost_lo + 1 : -1, counts);
    grouped_count(r.ghost_hi >= 0 ? 2 * r.ghost_hi + 1 : -1, counts);
}

__global__ void __launch_bounds__(kThreads) k_exchange_split(size_t n, const float4* __restrict__ posm,
                                                             const float4* __restrict__ velid, SlabCuts sc, int axis,
                                                             float ref_inv_cell, int layers, int me, float4* __restrict__ posm_out,
                                                             float4* __restrict__ velid_out, float4* __restrict__ rec,
                                                             ExchangeOffsets off, unsigned int* __restrict__ cursors) {
    const size_t s = (size_t)blockIdx.x * kThreads + threadIdx.x;
    Route r = {-1, -1, -1};
    float4 p, v;
    if (s < n) {
        v = velid[s];
        if (!(__float_as_uint(v.w) & kGhostBit)) {
            p = posm[s];
            r = route_of(sc, axis_scaled(p, axis, ref_inv_cell), layers);
        }
    }
    const int k0 = r.owner >= 0 ? 2 * r.owner : -1;
    const int k1 = r.ghost_lo >= 0 ? 2 * r.ghost_lo + 1 : -1;
    const int k2 = r.ghost_hi >= 0 ? 2 * r.ghost_hi + 1 : -1;
    const unsigned s0 = grouped_slot(k0, cursors);
    const unsigned s1 = grouped_slot(k1, cursors);
    const unsigned s2 = grouped_slot(k2, cursors);
    if (k0 < 0) return;
    if (r.owner == me) {
        posm_out[s0] = p;
        velid_out[s0] = v;
    } else {
        const size_t q = (size_t)off.start[k0] + s0;
        rec[2 * q] = p;
        rec[2 * q + 1] = v;
    }
    float4 g = v;
    g.w = __uint_as_float(__float_as_uint(v.w) | kGhostBit);
    if (k1 >= 0) { const size_t q = (size_t)off.start[k1] + s1; rec[2 * q] = p; rec[2 * q + 1] = g; }
    if (k2 >= 0) { const size_t q = (size_t)off.start[k2] + s2; rec[2 * q] = p; rec[2 * q + 1] = g; }
}

// records keep the ghost flag they arrive with
__global__ void __launch_bounds__(kThreads) k_slab_append_asis(size_t count, const float4* __restrict__ rec,
                                                               float4* __restrict__ posm, float4* __restrict__ velid) {
    const size_t k = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (k >= count) return;
    posm[k] = rec[2 * k];
    velid[k] = rec[2 * k + 1];
}

}  // namespace

int launch_slab_append(size_t count, const float4* rec, bool ghost, float4* posm, float4* velid, cudaStream_t st) {
    if (count == 0) return 0;
    k_slab_append<<<blocks_for(count), kThreads, 0, st>>>(count, rec, ghost ? kGhostBit : 0u, posm, velid);
    return 1;
}

int launch_slab_export(size_t n, const float4* posm, const float4* velid, const float2* rho_p, const float4* acc,
                       uint32_t* ids, float* pos3, float* vel3, float* rho, float* P, float* acc3, unsigned int* cursor,
                       cudaStream_t st) {
    if (n == 0) return 0;
    k_slab_export<<<blocks_for(n), kThreads, 0, st>>>(n, posm, velid, rho_p, acc, ids, pos3, vel3, rho, P, acc3, cursor);
    return 1;
}

int launch_pack_upload_ids(size_t n, const float* d_pos3, const float* d_vel3, const float* d_mass, const uint32_t* d_ids,
                           float default_mass, float4* posm, float4* velid, cudaStream_t st) {
    if (n == 0) return 0;
    k_pack_upload_ids<<<blocks_for(n), kThreads, 0, st>>>(n, d_pos3, d_vel3, d_mass, d_ids, default_mass, posm, velid);
    return 1;
}

int launch_exchange_count(size_t n, const float4* posm, const float4* velid, const SlabCuts& sc, int axis, float ref_inv_cell,
                          int layers, unsigned int* counts, cudaStream_t st) {
    if (n == 0) return 0;
    k_exchange_count<<<blocks_for(n), kThreads, 0, st>>>(n, posm, velid, sc, axis, ref_inv_cell, layers, counts);
    return 1;
}

int launch_exchange_split(size_t n, const float4* posm, const float4* velid, const SlabCuts& sc, int axis, float ref_inv_cell,
                          int layers, int me, float4* posm_out, float4* velid_out, float4* rec, const ExchangeOffsets& off,
                          unsigned int* cursors, cudaStream_t st) {
    if (n == 0) return 0;
    k_exchange_split<<<blocks_for(n), kThreads, 0, st>>>(n, posm, velid, sc, axis, ref_inv_cell, layers, me, posm_out, velid_out,
                                                          rec, off, cursors);
    return 1;
}

int launch_slab_append_asis(size_t count, const float4* rec, float4* posm, float4* velid, cudaStream_t st) {
    if (count == 0) return 0;
    k_slab_append_asis<<<blocks_for(count), kThreads, 0, st>>>(count, rec, posm, velid);
    return 1;
}

}  // namespace sphb
