// Internal declarations shared by the CUDA translation units of libsphb.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "sphb.h"

namespace sphb {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs; grids of the persistent kernels are sized from this

// ---- cell grid ----------------------------------------------------------------------------------
// The reference hashes cell coordinates c = floor(p * inv_cell) into a 63-bit key
// ((cx & 0x1FFFFF) << 42 | (cy & 0x1FFFFF) << 21 | (cz & 0x1FFFFF), spatial_hash.h:20-27).  Sorting
// by that key is lexicographic in the MASKED coordinates, i.e. on each axis the non-negative cells
// come first (ascending) and the negative ones after them (ascending).  `rank` reproduces exactly
// that order inside the bounding cell box [lo, hi], so the dense linear cell index
//     cell = (rank_x * ext_y + rank_y) * ext_z + rank_z
// is strictly monotone in the reference key and a sort by it yields the reference-key order.
//
// The per-axis fields are indexed by KEY POSITION k (0 = most significant, 2 = fastest-varying = the axis the
// contiguous cell columns run along); perm[k] is the physical axis (0 = x, 1 = y, 2 = z) at that position.
// Strict mode and the reference-order debug keys use the identity permutation with the masked-key rank
// (monotone = 0).  The fast-mode layout is free to differ from the reference order (the reference-order
// permutation is then produced by a separate debug sort): it uses monotone ranks (rank = c - lo, so a cell
// column is ONE contiguous run) and, in slab mode, puts the slab axis first so that the ghost layers are
// contiguous blocks of the sorted arrays (whole warps of ghosts exit the force pass at once).
struct GridDesc {
    float inv_cell;        // 1 / (internal cell size) — internal cell = neighbor_search_radius / refine
    float ref_inv_cell;    // 1 / neighbor_search_radius: the reference's hash cell (spatial_hash.h:63-66), for the 63-bit keys
    int lo[3], hi[3];      // inclusive cell-coordinate box covering every position that can occur
    int pos_lo[3];         // max(lo, 0)
    int npos[3];           // number of non-negative cells on the axis
    int ext[3];            // hi - lo + 1
    int perm[3];           // physical axis at key position k
    int monotone;          // 1: rank = c - lo ; 0: the reference's masked-key order (non-negative cells first)
    int pad;               // fast-mode layout: empty guard cells on every side of the populated box [lo + pad, hi - pad]
                           // (>= the walk radius, so the stencil walks need no range checks); 0 otherwise
    uint32_t ncells;
    int id_bits;           // bits of the particle id inside the composite sort key
    int cell_bits;
};

__host__ __device__ inline int grid_rank(const GridDesc& g, int k, int c) {
    if (g.monotone) return c - g.lo[k];
    return c >= 0 ? c - g.pos_lo[k] : c - g.lo[k] + g.npos[k];
}

// ---- physics constants, precomputed on the host in the reference's own fp32 expression order ----
struct PairConsts {
    float h;            // smoothing length (CubicSplineKernel::h_)
    float h_sq;         // h*h (kernels.cpp:13)
    float sigma;        // 1/(pi*h*h*h) (kernels.cpp:27)
    float r2;           // nsr*nsr (sph_engine.cpp:347)
    float w0;           // W(0) = sigma*(2/3 - 0 + 0) (kernels.cpp:60)
    float rest_density, gas_constant, viscosity, gravity;
    // fast-mode folded constants
    float inv_h;        // 1/h
    float sig_h;        // sigma / h
    float sig_h2;       // sigma / h_sq
    float neg_zero;     // -0.0f, deliberately a RUN-TIME value: see f2_sq_exact in pair.cu
};

struct IntegrateConsts {
    float damping;
    float xmin, xmax, ymin, ymax, zmin, zmax;
    float cfl, h, timestep;
};

// Small block of device-resident scalars (one allocation, zero-initialised).
struct DeviceScalars {
    float dt;             // dt of the step being executed
    float time;           // SPHEngine::current_time_, accumulated on the device in fp32
    float a0[3];          // accelerations_[0] of the last step (compute_cfl_timestep, sph_engine.cpp:330)
    unsigned int max_v2_bits;   // max |v|^2 over the current velocities, as float bits (non-negative)
    unsigned int max_neighbors; // running max of neighbour-list length, self included
    unsigned int error_flags;   // bit0: a particle fell outside the cell box
    double sum_rho, kinetic;    // diagnostics scratch
    unsigned int diag_max_v2_bits;
    unsigned int a0_fresh;      // set by k_integrate when it advanced particle id 0 (slab mode: who owns a0)
};

// ---- launch wrappers (each returns the number of kernels it enqueued) ---------------------------
struct SortBuffers {
    uint64_t* keys[2];
    uint32_t* vals[2];
    uint32_t* counts;     // 256 * ntiles digit counts, scanned in place
    uint32_t* block_sums; // scratch of the device-wide scan
    size_t counts_cap;    // elements available in counts
    size_t block_sums_cap;
};

constexpr int kSortTile = 2048;   // keys per CTA tile of the radix sort
constexpr int kScanTile = 4096;   // elements per CTA of the device-wide scan

// LSD radix sort of (key, val) pairs by the low `bits` bits of key, 8 bits per pass, stable.
// Input in buffers [0]; returns the index (0/1) of the buffer pair holding the result in *out_buf.
int launch_radix_sort(const SortBuffers& sb, size_t n, int bits, int* out_buf, cudaStream_t st);
// Stable sort by the key bits [first_bit, first_bit + bits), one scatter kernel per 8-bit digit with decoupled
// look-back; the first pass generates vals = slot itself (the caller need not fill vals[0]).  scratch must hold
// onesweep_scratch_bytes(n_max, bits_max).
size_t onesweep_scratch_bytes(size_t n_max, int bits_max);
int launch_radix_sort_onesweep(const SortBuffers& sb, size_t n, int first_bit, int bits, int* out_buf, void* scratch,
                               cudaStream_t st);

// out[i] = sum(in[0..i-1]) (exclusive) or out[i] = max(in[0..i]) (inclusive); in == out allowed.
int launch_scan_sum_exclusive(uint32_t* data, size_t n, uint32_t* block_sums, cudaStream_t st);
int launch_scan_max_inclusive(uint32_t* data, size_t n, uint32_t* block_sums, cudaStream_t st);

int launch_pack_upload(size_t n, const float* d_pos3, const float* d_vel3, const float* d_mass, float default_mass,
                       float4* posm, float4* velid, DeviceScalars* sc, cudaStream_t st);
// refkeys (63-bit reference keys) and ckeys (composite keys on the coarse reference grid gc) are optional
// debug outputs.
int launch_cell_keys(size_t n, const float4* posm, const float4* velid, GridDesc g, uint64_t* keys, uint32_t* vals,
                     uint64_t* refkeys_or_null, uint64_t* ckeys_or_null, GridDesc gc, DeviceScalars* sc, cudaStream_t st);
// cell_start has ncells + 1 entries; block_sums is scan scratch.
int launch_cell_table(size_t n, const uint64_t* sorted_keys, GridDesc g, uint32_t* cell_start, uint32_t* block_sums,
                      cudaStream_t st);
// sorted_keys need only be sorted by their cell field (key >> id_bits): the reorder kernel ranks every particle among
// the particles of its cell by full key (= by id) and writes it to cell_start[cell] + rank, so the final layout is the
// (cell, id) order whether or not the sort looked at the id bits.
int launch_reorder(size_t n, const uint64_t* sorted_keys, int id_bits, const uint32_t* cell_start, const uint32_t* sorted_vals,
                   const float4* posm_in, const float4* velid_in, const uint64_t* refkeys_in, float4* posm_out, float4* velid_out,
                   uint64_t* refkeys_out, float4* pp2_out, cudaStream_t st);

// variant 2: everything the force pass needs from a neighbour, in one 32-byte record (A = m / (2 rho), B = A * P)
struct __align__(32) ForceRec {
    float x, y, z, A;
    float vx, vy, vz, B;
};

struct PairArgs {
    size_t n;
    const float4* posm;        // x, y, z, mass   (cell-sorted)
    const float4* velid;       // vx, vy, vz, id bits
    const uint32_t* cell_start;
    float2* rho_p;             // density, pressure
    float4* fa;                // force-pass staging A (fast: x,y,z,m/(2 rho) ; strict: x,y,z,m)
    float4* fb;                // force-pass staging B (fast: v, A*P ; strict: v, rho)
    // pair-interleaved mirrors for the packed-f32x2 kernels: record k = particles 2k, 2k+1 as
    // {x0,x1,y0,y1 | z0,z1,w0,w1} (two float4)
    const float4* pp2;         // position + mass pairs (written by the reorder kernel)
    float4* fa2;               // x, y, z, A' = (sigma/h) m / (2 rho)
    float4* fb2;               // vx, vy, vz, B' = A' P
    float4* acc;               // ax, ay, az, (unused)
    // variant 2: accepted-neighbour bitmasks handed from the density pass to the force pass, column-major:
    // masks[col * mask_stride + slot], col = 0 .. mask_cols(R)-1 are the (2R+1)^2 cell columns in walk order (bit k =
    // k-th candidate of the column run), col = mask_cols(R) is {overflow flag, neighbour count}
    void* masks;               // uint2 per (column, particle) for R <= 3, uint32 for R = 4 (mask_words())
    size_t mask_stride;
    ForceRec* fab;             // variant 2: force-pass records, written by the density pass
    uint32_t* nbr_count;       // optional, per sorted slot
    DeviceScalars* sc;
    GridDesc grid;
    PairConsts k;
    int walk_radius;
    int strict;
    int variant;
    // slab mode (slab_axis >= 0): density is evaluated for particles whose reference cell on the axis lies in
    // [rho_lo, rho_hi) (owned + one halo layer); force only for owned particles (id word bit 31 clear)
    int slab_axis;
    int rho_lo, rho_hi;
};
int launch_density(const PairArgs& a, cudaStream_t st);
int launch_force(const PairArgs& a, cudaStream_t st);
// variant 2 walks (2R+1)^2 columns of 2R+1 cells, R = walk_radius * grid_refine in [kMaskMinRadius, kMaskMaxRadius]
constexpr int kMaskMinRadius = 2, kMaskMaxRadius = 6;
constexpr float kRefinedCellScale = 0.9990234375f;   // 1 - 2^-10: refined cells are this much larger than nsr / refine (make_grid)
constexpr int mask_cols(int R) { return (2 * R + 1) * (2 * R + 1); }
#ifndef SPHB_MASK_W4
#define SPHB_MASK_W4 1
#endif
constexpr int mask_words(int R) { return R >= 4 ? SPHB_MASK_W4 : 2; }   // 32-bit words per column mask (a column holds ~(2R+1)/R^3 of a coarse cell)
int stencil_reach_table(int R, signed char* out);   // host: copies the (2R+1)^2 column reaches, returns their number or -1
int launch_density_mask(const PairArgs& a, cudaStream_t st);
int launch_force_mask(const PairArgs& a, cudaStream_t st);

int launch_set_dt(DeviceScalars* sc, float dt, cudaStream_t st);
int launch_cfl_dt(DeviceScalars* sc, IntegrateConsts ic, cudaStream_t st);
int launch_integrate(size_t n, float4* posm, float4* velid, const float4* acc, IntegrateConsts ic, DeviceScalars* sc,
                     int* d_box_or_null, cudaStream_t st);
int launch_max_speed(size_t n, const float4* velid, DeviceScalars* sc, cudaStream_t st);

int launch_unpermute(size_t n, const float4* posm, const float4* velid, const float2* rho_p, const float4* acc,
                     const uint64_t* refkeys, const uint32_t* nbr_count, float* pos3, float* vel3, float* rho, float* P,
                     float* acc3, uint64_t* keys, uint32_t* perm, uint32_t* counts, cudaStream_t st);
int launch_diagnostics(size_t n, const float4* posm, const float4* velid, const float2* rho_p, DeviceScalars* sc,
                       cudaStream_t st);

// ---- slab decomposition (slab.cu) ----------------------------------------------------------------
constexpr int kMaxRanks = 64;
struct SlabCuts { int nranks; int cuts[kMaxRanks + 1]; };       // rank d owns reference cells [cuts[d], cuts[d+1])
struct ExchangeOffsets { unsigned int start[2 * kMaxRanks]; };   // one-round exchange: key 2r = owned by r, 2r+1 = ghost for r
int launch_exchange_count(size_t n, const float4* posm, const float4* velid, const SlabCuts& sc, int axis, float ref_inv_cell,
                          int layers, unsigned int* counts, cudaStream_t st);
int launch_exchange_split(size_t n, const float4* posm, const float4* velid, const SlabCuts& sc, int axis, float ref_inv_cell,
                          int layers, int me, float4* posm_out, float4* velid_out, float4* rec, const ExchangeOffsets& off,
                          unsigned int* cursors, cudaStream_t st);
int launch_slab_append_asis(size_t count, const float4* rec, float4* posm, float4* velid, cudaStream_t st);
int launch_slab_append(size_t count, const float4* rec, bool ghost, float4* posm, float4* velid, cudaStream_t st);
int launch_slab_export(size_t n, const float4* posm, const float4* velid, const float2* rho_p, const float4* acc,
                       uint32_t* ids, float* pos3, float* vel3, float* rho, float* P, float* acc3, unsigned int* cursor,
                       cudaStream_t st);
int launch_pack_upload_ids(size_t n, const float* d_pos3, const float* d_vel3, const float* d_mass, const uint32_t* d_ids,
                           float default_mass, float4* posm, float4* velid, cudaStream_t st);

}  // namespace sphb
