// Spatial-hash stage on the device: cell keys, dense cell table, SoA reorder, un-permute.
//
// Replaces ParticleSystem::get_positions + SpatialHash::build (reference src/particle.cpp:68-75,
// src/spatial_hash.cpp:15-25) and the id-order getters (src/sph_engine.h:133-136).
// All kernels are HBM-bound streaming passes over float4 SoA columns.
#include "sphb_internal.cuh"

namespace sphb {

namespace {

constexpr int kThreads = 256;

inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// (int)floorf(p * inv_cell), reference spatial_hash.h:30-36.  __float2int_rd is floor + convert and,
// unlike the CPU cast, is defined for NaN/out-of-range (0 / saturation) — SURVEY Q22.
__device__ __forceinline__ int cell_coord(float p, float inv_cell) { return __float2int_rd(__fmul_rn(p, inv_cell)); }

__device__ __forceinline__ void block_max_v2(float v2, DeviceScalars* sc) {
    // |v|^2 >= 0, so its float bit pattern orders like an unsigned integer; NaN is dropped the way
    // std::max(max_velocity, velocity) drops it (sph_engine.cpp:318).
    unsigned bits = (v2 == v2) ? __float_as_uint(v2) : 0u;
    bits = __reduce_max_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0 && bits > *(volatile unsigned int*)&sc->max_v2_bits) atomicMax(&sc->max_v2_bits, bits);
}

// compute_cfl_timestep, reference sph_engine.cpp:312-333: min(CFL h / (max|v| + 1e-6), CFL sqrt(h / (|a_0| + 1e-6)), timestep);
// the force criterion reads accelerations_[0] only.  std::min({a, b, c}) returns the first of the smallest.
__device__ __forceinline__ float cfl_timestep(const DeviceScalars* sc, const IntegrateConsts& ic) {
    const float max_velocity = __fsqrt_rn(__uint_as_float(sc->max_v2_bits));
    const float dt_cfl = __fdiv_rn(__fmul_rn(ic.cfl, ic.h), __fadd_rn(max_velocity, 1e-6f));
    const float a0 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(sc->a0[0], sc->a0[0]), __fmul_rn(sc->a0[1], sc->a0[1])),
                                          __fmul_rn(sc->a0[2], sc->a0[2])));
    const float dt_force = __fmul_rn(ic.cfl, __fsqrt_rn(__fdiv_rn(ic.h, __fadd_rn(a0, 1e-6f))));
    float m = dt_cfl;
    if (dt_force < m) m = dt_force;
    if (ic.timestep < m) m = ic.timestep;
    return m;
}

// sphb_cfl_timestep: evaluates the rule without consuming the running maximum
__global__ void k_cfl_probe(DeviceScalars* sc, IntegrateConsts ic) { sc->dt = cfl_timestep(sc, ic); }

__global__ void __launch_bounds__(kThreads) k_pack_upload(size_t n, const float* __restrict__ pos3,
                                                          const float* __restrict__ vel3, const float* __restrict__ mass,
                                                          float default_mass, float4* __restrict__ posm,
                                                          float4* __restrict__ velid, DeviceScalars* sc) {
    size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    float v2 = 0.0f;
    if (i < n) {
        float4 p, v;
        p.x = pos3[3 * i]; p.y = pos3[3 * i + 1]; p.z = pos3[3 * i + 2];
        p.w = mass ? mass[i] : default_mass;
        if (vel3) { v.x = vel3[3 * i]; v.y = vel3[3 * i + 1]; v.z = vel3[3 * i + 2]; }
        else { v.x = 0.0f; v.y = 0.0f; v.z = 0.0f; }
        v.w = __uint_as_float((unsigned)i);
        posm[i] = p;
        velid[i] = v;
        v2 = __fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z));
    }
    block_max_v2(v2, sc);
}

__device__ __forceinline__ uint32_t dense_cell(const GridDesc& g, const float4& p, bool* outside) {
    uint32_t cell = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int a = g.perm[k];
        int c = cell_coord(a == 0 ? p.x : (a == 1 ? p.y : p.z), g.inv_cell);
        if (c < g.lo[k] + g.pad) { c = g.lo[k] + g.pad; *outside = true; }
        if (c > g.hi[k] - g.pad) { c = g.hi[k] - g.pad; *outside = true; }
        cell = cell * (uint32_t)g.ext[k] + (uint32_t)grid_rank(g, k, c);
    }
    return cell;
}

// ---- counting sort by cell (round 2; replaces the radix sort of (cell, id) keys on the per-step path) ----------------
// SpatialHash::build (reference src/spatial_hash.cpp:15-25) pushes every particle index into its cell's vector.  The
// device equivalent is a counting sort over the dense cell table:
//   k_cell_count   cell of every particle + one atomic per particle on the cell's counter (the returned ticket is
//                  the particle's provisional rank inside its cell), and the step's dt (fixed, or the CFL rule)
//   scan           exclusive prefix sum of the counters, in place -> cell_start (one pass, decoupled look-back)
//   k_cell_scatter slot_src[cell_start[cell] + ticket] = particle
//   k_reorder      gathers the records into cell order; inside a cell the ascending-id order of the reference's
//                  per-cell vectors (which fixes the summation order) is restored by ranking the cell's members
// 150 B of traffic per particle and 4 kernels instead of ~250 B and 10 kernels for the radix-sort path.
__global__ void __launch_bounds__(kThreads) k_cell_count(size_t n, const float4* __restrict__ posm,
                                                         const float4* __restrict__ velid, GridDesc g,
                                                         uint2* __restrict__ cell_ticket, uint32_t* __restrict__ cell_cnt,
                                                         uint64_t* __restrict__ refkeys, uint64_t* __restrict__ ckeys, GridDesc gc,
                                                         DeviceScalars* sc, float dt_fixed, IntegrateConsts ic) {
    size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i == 0) {
        // dt of this step: the caller's, or compute_cfl_timestep (sph_engine.cpp:312-333) from the running max |v|^2 and
        // accelerations_[0] of the previous step.  max |v| = sqrt(max |v|^2) because correctly rounded sqrt is monotone.
        float dt = dt_fixed;
        if (dt_fixed <= 0.0f) dt = cfl_timestep(sc, ic);
        sc->dt = dt;
        sc->max_v2_bits = 0u;   // consumed; k_integrate of this step accumulates the next value
    }
    if (i >= n) return;
    const float4 p = posm[i];
    if (refkeys) {
        // SpatialHash::get_grid_coords + hash_position, reference spatial_hash.h:20-36
        const int cx = cell_coord(p.x, g.ref_inv_cell), cy = cell_coord(p.y, g.ref_inv_cell), cz = cell_coord(p.z, g.ref_inv_cell);
        refkeys[i] = ((uint64_t)(cx & 0x1FFFFF) << 42) | ((uint64_t)(cy & 0x1FFFFF) << 21) | (uint64_t)(cz & 0x1FFFFF);
    }
    bool outside = false;
    const uint32_t cell = dense_cell(g, p, &outside);
    if (outside) atomicOr(&sc->error_flags, 1u);
    cell_ticket[i] = make_uint2(cell, atomicAdd(&cell_cnt[cell], 1u));
    if (ckeys) {   // debug: (reference cell, id) composite — sorting it yields the reference-order permutation
        const unsigned id = __float_as_uint(velid[i].w) & 0x7FFFFFFFu;   // bit 31 marks halo copies in slab mode
        bool o2 = false;
        ckeys[i] = ((uint64_t)dense_cell(gc, p, &o2) << gc.id_bits) | (uint64_t)id;
    }
}

__global__ void __launch_bounds__(kThreads) k_cell_scatter(size_t n, const uint2* __restrict__ cell_ticket,
                                                           const uint32_t* __restrict__ cell_start,
                                                           uint32_t* __restrict__ slot_src) {
    size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const uint2 ct = cell_ticket[i];
    slot_src[cell_start[ct.x] + ct.y] = (uint32_t)i;
}

// Tickets are handed out in arrival order; the id order inside a cell — the order of the reference's per-cell vectors
// (spatial_hash.cpp:19-24) — is restored here: a particle's final slot is cell_start[cell] + (number of members of its
// cell with a smaller id).  Cells hold ~1 (refined fast grid) to ~64 (strict) particles; collapsed cells cost
// occupancy^2 reads, the same order as the pair passes themselves.
__global__ void __launch_bounds__(kThreads) k_reorder(size_t n, const uint32_t* __restrict__ slot_src,
                                                      const uint2* __restrict__ cell_ticket,
                                                      const uint32_t* __restrict__ cell_start,
                                                      const float4* __restrict__ posm_in, const float4* __restrict__ velid_in,
                                                      const uint64_t* __restrict__ refkeys_in, float4* __restrict__ posm_out,
                                                      float4* __restrict__ velid_out, uint64_t* __restrict__ refkeys_out) {
    size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x;
    // slot n is a massless sentinel: the pair kernels read candidates two at a time and may touch one slot past the
    // last run (its contribution is masked, but it must be finite); the buffers hold capacity + 4 records
    if (t == 0) posm_out[n] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (t >= n) return;
    const uint32_t src = slot_src[t];
    const uint32_t cell = cell_ticket[src].x;
    const uint32_t cs = cell_start[cell], ce = cell_start[cell + 1];
    const float4 v = velid_in[src];
    uint32_t rank = 0;
    if (ce - cs > 1u) {
        const uint32_t id = __float_as_uint(v.w) & 0x7FFFFFFFu;
        for (uint32_t u = cs; u < ce; ++u)
            rank += ((__float_as_uint(__ldg(&velid_in[slot_src[u]].w)) & 0x7FFFFFFFu) < id) ? 1u : 0u;
    }
    const size_t s = (size_t)cs + rank;
    posm_out[s] = posm_in[src];
    velid_out[s] = v;
    if (refkeys_in) refkeys_out[s] = refkeys_in[src];
}

__global__ void __launch_bounds__(kThreads) k_unpermute(size_t n, const float4* __restrict__ posm,
                                                        const float4* __restrict__ velid, const float2* __restrict__ rho_p,
                                                        const float4* __restrict__ acc, const uint64_t* __restrict__ refkeys,
                                                        const uint32_t* __restrict__ nbr_count, float* __restrict__ pos3,
                                                        float* __restrict__ vel3, float* __restrict__ rho, float* __restrict__ P,
                                                        float* __restrict__ acc3, uint64_t* __restrict__ keys,
                                                        uint32_t* __restrict__ perm, uint32_t* __restrict__ counts) {
    size_t s = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (s >= n) return;
    float4 v = velid[s];
    size_t id = __float_as_uint(v.w);
    if (pos3) { float4 p = posm[s]; pos3[3 * id] = p.x; pos3[3 * id + 1] = p.y; pos3[3 * id + 2] = p.z; }
    if (vel3) { vel3[3 * id] = v.x; vel3[3 * id + 1] = v.y; vel3[3 * id + 2] = v.z; }
    if (rho || P) { float2 rp = rho_p[s]; if (rho) rho[id] = rp.x; if (P) P[id] = rp.y; }
    if (acc3) { float4 a = acc[s]; acc3[3 * id] = a.x; acc3[3 * id + 1] = a.y; acc3[3 * id + 2] = a.z; }
    if (keys) keys[id] = refkeys[s];
    if (perm) perm[s] = (uint32_t)id;
    if (counts) counts[id] = nbr_count[s];
}

// The renderer's per-instance record (reference Renderer::update_particle_data, src/renderer.cpp:279-312): position (3),
// velocity (3), colour (3) per particle in insertion order.  colours: per-id RGB uploaded once (Particle::color is never
// touched by the physics, quirk Q8) or NULL for one default colour.
__global__ void __launch_bounds__(kThreads) k_export_instances(size_t n, const float4* __restrict__ posm,
                                                               const float4* __restrict__ velid, const float* __restrict__ colors,
                                                               float3 default_color, float* __restrict__ out9) {
    const size_t s = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (s >= n) return;
    const float4 v = velid[s], p = posm[s];
    const size_t id = __float_as_uint(v.w) & 0x7FFFFFFFu;
    float* o = out9 + 9 * id;
    o[0] = p.x; o[1] = p.y; o[2] = p.z;
    o[3] = v.x; o[4] = v.y; o[5] = v.z;
    if (colors) { o[6] = colors[3 * id]; o[7] = colors[3 * id + 1]; o[8] = colors[3 * id + 2]; }
    else { o[6] = default_color.x; o[7] = default_color.y; o[8] = default_color.z; }
}

__global__ void __launch_bounds__(kThreads) k_diagnostics(size_t n, const float4* __restrict__ posm,
                                                          const float4* __restrict__ velid, const float2* __restrict__ rho_p,
                                                          DeviceScalars* sc) {
    __shared__ double s_rho[kThreads / 32], s_ke[kThreads / 32];
    double rho = 0.0, ke = 0.0;
    unsigned vbits = 0;
    for (size_t s = (size_t)blockIdx.x * kThreads + threadIdx.x; s < n; s += (size_t)gridDim.x * kThreads) {
        float4 v = velid[s];
        if (__float_as_uint(v.w) & 0x80000000u) continue;   // slab mode: halo copies belong to another context
        float m = posm[s].w;
        float v2 = __fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z));
        rho += (double)rho_p[s].x;
        ke += 0.5 * (double)m * (double)v2;
        if (v2 == v2) vbits = max(vbits, __float_as_uint(v2));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        rho += __shfl_down_sync(0xffffffffu, rho, d);
        ke += __shfl_down_sync(0xffffffffu, ke, d);
    }
    vbits = __reduce_max_sync(0xffffffffu, vbits);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_rho[w] = rho; s_ke[w] = ke; if (vbits) atomicMax(&sc->diag_max_v2_bits, vbits); }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < kThreads / 32; ++i) { a += s_rho[i]; b += s_ke[i]; }
        atomicAdd(&sc->sum_rho, a);
        atomicAdd(&sc->kinetic, b);
    }
}

__global__ void __launch_bounds__(kThreads) k_max_speed(size_t n, const float4* __restrict__ velid, DeviceScalars* sc) {
    size_t s = (size_t)blockIdx.x * kThreads + threadIdx.x;
    float v2 = 0.0f;
    if (s < n) {
        float4 v = velid[s];
        v2 = __fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z));
    }
    block_max_v2(v2, sc);
}

// Signed-int image of a float that orders like the float (for atomicMin/atomicMax).
__device__ __forceinline__ int ordered_int(float f) {
    int s = __float_as_int(f);
    return s >= 0 ? s : s ^ 0x7FFFFFFF;
}

__global__ void __launch_bounds__(kThreads) k_bbox(size_t n, const float4* __restrict__ posm, int* box) {
    int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF};
    int hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const float4 p = posm[i];
        const float c[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (fabsf(c[a]) <= 3.4028235e38f) {   // finite only: NaN/inf are clamped into the box by k_cell_keys
                const int o = ordered_int(c[a]);
                lo[a] = min(lo[a], o);
                hi[a] = max(hi[a], o);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
        hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&box[a], lo[a]);
            atomicMax(&box[3 + a], hi[a]);
        }
    }
}

// Array-of-structs input (e.g. the reference's 76-byte sph::Particle, particle.h:17-49) → float4 SoA.
__global__ void __launch_bounds__(kThreads) k_unpack_strided(size_t n, const unsigned char* __restrict__ base, size_t stride,
                                                             size_t off_pos, size_t off_vel, size_t off_mass,
                                                             float4* __restrict__ posm, float4* __restrict__ velid,
                                                             DeviceScalars* sc) {
    size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    float v2 = 0.0f;
    if (i < n) {
        const unsigned char* rec = base + i * stride;
        const float* p = reinterpret_cast<const float*>(rec + off_pos);
        const float* v = reinterpret_cast<const float*>(rec + off_vel);
        const float m = *reinterpret_cast<const float*>(rec + off_mass);
        posm[i] = make_float4(p[0], p[1], p[2], m);
        velid[i] = make_float4(v[0], v[1], v[2], __uint_as_float((unsigned)i));
        v2 = __fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2]));
    }
    block_max_v2(v2, sc);
}

}  // namespace

int launch_pack_upload(size_t n, const float* d_pos3, const float* d_vel3, const float* d_mass, float default_mass,
                       float4* posm, float4* velid, DeviceScalars* sc, cudaStream_t st) {
    if (n == 0) return 0;
    k_pack_upload<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, d_pos3, d_vel3, d_mass, default_mass, posm, velid, sc);
    return 1;
}

int launch_cell_count(size_t n, const float4* posm, const float4* velid, GridDesc g, uint2* cell_ticket, uint32_t* cell_cnt,
                      uint64_t* refkeys_or_null, uint64_t* ckeys_or_null, GridDesc gc, DeviceScalars* sc, float dt_fixed,
                      IntegrateConsts ic, cudaStream_t st) {
    k_cell_count<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, posm, velid, g, cell_ticket, cell_cnt, refkeys_or_null,
                                                                ckeys_or_null, gc, sc, dt_fixed, ic);
    return 1;
}

int launch_cell_scatter(size_t n, const uint2* cell_ticket, const uint32_t* cell_start, uint32_t* slot_src, cudaStream_t st) {
    k_cell_scatter<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, cell_ticket, cell_start, slot_src);
    return 1;
}

int launch_reorder(size_t n, const uint32_t* slot_src, const uint2* cell_ticket, const uint32_t* cell_start,
                   const float4* posm_in, const float4* velid_in, const uint64_t* refkeys_in, float4* posm_out, float4* velid_out,
                   uint64_t* refkeys_out, cudaStream_t st) {
    k_reorder<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, slot_src, cell_ticket, cell_start, posm_in, velid_in, refkeys_in,
                                                             posm_out, velid_out, refkeys_out);
    return 1;
}

int launch_cfl_probe(DeviceScalars* sc, IntegrateConsts ic, cudaStream_t st) {
    k_cfl_probe<<<1, 1, 0, st>>>(sc, ic);
    return 1;
}

int launch_unpermute(size_t n, const float4* posm, const float4* velid, const float2* rho_p, const float4* acc,
                     const uint64_t* refkeys, const uint32_t* nbr_count, float* pos3, float* vel3, float* rho, float* P,
                     float* acc3, uint64_t* keys, uint32_t* perm, uint32_t* counts, cudaStream_t st) {
    if (n == 0) return 0;
    k_unpermute<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, posm, velid, rho_p, acc, refkeys, nbr_count, pos3, vel3,
                                                               rho, P, acc3, keys, perm, counts);
    return 1;
}

int launch_export_instances(size_t n, const float4* posm, const float4* velid, const float* colors, const float default_color[3],
                            float* out9, cudaStream_t st) {
    if (n == 0) return 0;
    k_export_instances<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, posm, velid, colors,
                                                                     make_float3(default_color[0], default_color[1], default_color[2]), out9);
    return 1;
}

int launch_diagnostics(size_t n, const float4* posm, const float4* velid, const float2* rho_p, DeviceScalars* sc,
                       cudaStream_t st) {
    if (n == 0) return 0;
    unsigned nb = blocks_for(n, kThreads);
    if (nb > (unsigned)(kSMs * 8)) nb = kSMs * 8;
    k_diagnostics<<<nb, kThreads, 0, st>>>(n, posm, velid, rho_p, sc);
    return 1;
}

int launch_max_speed(size_t n, const float4* velid, DeviceScalars* sc, cudaStream_t st) {
    if (n == 0) return 0;
    k_max_speed<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, velid, sc);
    return 1;
}

int launch_bbox(size_t n, const float4* posm, int* d_box, cudaStream_t st) {
    if (n == 0) return 0;
    unsigned nb = blocks_for(n, kThreads);
    if (nb > (unsigned)(kSMs * 8)) nb = kSMs * 8;
    k_bbox<<<nb, kThreads, 0, st>>>(n, posm, d_box);
    return 1;
}

// Six words from device memory into PINNED HOST memory, written by a kernel: a read-back that does not pass through a
// copy engine, so it never queues behind a bulk transfer of another stream (sphb_download_begin) on that engine.
namespace {
__global__ void k_box_to_host(const int* __restrict__ d_box, volatile int* h_box) {
    if (threadIdx.x < 6) h_box[threadIdx.x] = d_box[threadIdx.x];
}
}  // namespace
int launch_box_to_host(const int* d_box, int* h_box_pinned, cudaStream_t st) {
    k_box_to_host<<<1, 32, 0, st>>>(d_box, h_box_pinned);
    return 1;
}
namespace {
__global__ void k_word_to_host(const int* __restrict__ d_word, volatile int* h_word) { *h_word = *d_word; }
}  // namespace
int launch_word_to_host(const int* d_word, int* h_word_pinned, cudaStream_t st) {
    k_word_to_host<<<1, 1, 0, st>>>(d_word, h_word_pinned);
    return 1;
}
namespace {
__global__ void k_words_to_host(const uint32_t* __restrict__ d_src, volatile uint32_t* h_dst, unsigned nwords) {
    for (unsigned k = threadIdx.x; k < nwords; k += blockDim.x) h_dst[k] = d_src[k];
}
}  // namespace
int launch_words_to_host(const uint32_t* d_src, uint32_t* h_dst_pinned, unsigned nwords, cudaStream_t st) {
    if (nwords == 0) return 0;
    k_words_to_host<<<1, 256, 0, st>>>(d_src, h_dst_pinned, nwords);
    return 1;
}

int launch_unpack_strided(size_t n, const unsigned char* d_base, size_t stride, size_t off_pos, size_t off_vel, size_t off_mass,
                          float4* posm, float4* velid, DeviceScalars* sc, cudaStream_t st) {
    if (n == 0) return 0;
    k_unpack_strided<<<blocks_for(n, kThreads), kThreads, 0, st>>>(n, d_base, stride, off_pos, off_vel, off_mass, posm, velid, sc);
    return 1;
}

}  // namespace sphb
