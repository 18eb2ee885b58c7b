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
    float neg_zero;     // -0.0f, deliberately a RUN-TIME value: exact packed squares are formed as fma(d, d, -0) (pair_mask.cu)
    float r2_next;      // nextafterf(r2, +inf): d2 <= r2 <=> d2 - r2_next < 0 (sign-bit radius test of pair_mask.cu)
    // the other kernel classes behind create_kernel (SPHB_OPT_KERNEL_TYPE; tested-walk kernels of pair.cu only)
    float wnorm;        // WendlandC2Kernel::norm_factor_ = 21 / (2 pi h h h)   (kernels.cpp:168)
    float gssi;         // GaussianKernel::sigma_sq_inv_ = 1 / (h h)            (kernels.cpp:203)
    float gnorm;        // GaussianKernel::norm_ = 1 / pow(pi h h, 1.5)         (kernels.cpp:204)
};
// Kernel::W and friends (reference src/kernels.h:94-98 KernelType order)
constexpr int kKernelCubic = 0, kKernelWendlandC2 = 1, kKernelGaussian = 2;

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
struct SortBuffers {      // radix sort of the debug path (reference-order permutation)
    uint64_t* keys[2];
    uint32_t* vals[2];
};

constexpr int kSortTile = 2048;   // keys per CTA tile of the radix sort
constexpr int kScanTile = 4096;   // elements per CTA of the device-wide scan

// Stable sort by the key bits [first_bit, first_bit + bits), one scatter kernel per 8-bit digit with decoupled
// look-back; the first pass generates vals = slot itself (the caller need not fill vals[0]).  scratch must hold
// onesweep_scratch_bytes(n_max, bits_max).  Input in buffers [0]; *out_buf = index of the buffer pair holding the result.
size_t onesweep_scratch_bytes(size_t n_max, int bits_max);
int launch_radix_sort_onesweep(const SortBuffers& sb, size_t n, int first_bit, int bits, int* out_buf, void* scratch,
                               cudaStream_t st);

// In-place exclusive prefix sum of data[0..n) in one pass (decoupled look-back).  scratch: scan_scratch_bytes(n) bytes,
// ZERO on entry.
size_t scan_scratch_bytes(size_t n);
int launch_scan_exclusive(uint32_t* data, size_t n, void* scratch, cudaStream_t st);

int launch_pack_upload(size_t n, const float* d_pos3, const float* d_vel3, const float* d_mass, float default_mass,
                       float4* posm, float4* velid, DeviceScalars* sc, cudaStream_t st);
// Counting sort by cell (neighbor.cu).  launch_cell_count also sets the dt of the step (dt_fixed > 0: that value, else
// the CFL rule of compute_cfl_timestep); cell_cnt must be zero on entry.  refkeys (63-bit reference keys) and ckeys
// (composite keys on the coarse reference grid gc) are optional debug outputs.
int launch_cell_count(size_t n, const float4* posm, const float4* velid, GridDesc g, uint2* cell_ticket, uint32_t* cell_cnt,
                      uint64_t* refkeys_or_null, uint64_t* ckeys_or_null, GridDesc gc, DeviceScalars* sc, float dt_fixed,
                      IntegrateConsts ic, cudaStream_t st);
int launch_cell_scatter(size_t n, const uint2* cell_ticket, const uint32_t* cell_start, uint32_t* slot_src, cudaStream_t st);
// gathers the records into (cell, id) order: slot = cell_start[cell] + rank of the id among the cell's members
int launch_reorder(size_t n, const uint32_t* slot_src, const uint2* cell_ticket, const uint32_t* cell_start,
                   const float4* posm_in, const float4* velid_in, const uint64_t* refkeys_in, float4* posm_out, float4* velid_out,
                   uint64_t* refkeys_out, cudaStream_t st);

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
    float4* acc;               // ax, ay, az, (unused)
    // variant 2: accepted-neighbour bitmasks handed from the density pass to the force pass, row-major by column
    // group: masks[row * mask_stride + slot].  R >= 4 (pair_mask.cu): one uint32 per mirror pair of cell columns
    // (16 bits each) in walk order, empty corner groups skipped, last row = centre column + overflow flag.
    // R <= 3 (pair_mask_wide.cu): one uint2 per column, last row = {overflow flag, neighbour count}.
    void* masks;
    size_t mask_stride;
    ForceRec* fab;             // variant 2: force-pass records, written by the density pass
    uint32_t* nbr_count;       // optional, per sorted slot
    DeviceScalars* sc;
    GridDesc grid;
    PairConsts k;
    int walk_radius;
    int strict;
    int variant;
    int lanes;                 // variant 2, R >= 4: lanes sharing one particle (1 = pair_mask.cu / pair_stage.cu; 2, 4, 8 = pair_split.cu)
    int kernel_type;           // kKernelCubic / kKernelWendlandC2 / kKernelGaussian (the latter two: tested-walk kernels only)
    int mode;                  // variant 2, R >= 4: 0 = pair_mask.cu, 1 = pair_stage.cu, 2 = staged density + per-lane force (SPHB_OPT_PAIR_MODE)
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
size_t mask_bytes_per_slot(int R);                  // bytes of mask storage per unit of mask_stride
int stencil_reach_table(int R, signed char* out);   // host: copies the (2R+1)^2 column reaches, returns their number or -1
int launch_density_mask(const PairArgs& a, cudaStream_t st);
int launch_force_mask(const PairArgs& a, cudaStream_t st);
int launch_density_split(const PairArgs& a, cudaStream_t st);       // R >= 4, a.lanes lanes per particle (pair_split.cu)
int launch_force_split(const PairArgs& a, cudaStream_t st);
int launch_density_stage(const PairArgs& a, cudaStream_t st);       // R >= 4, shared-memory staged (pair_stage.cu); -1: not configurable
int launch_force_stage(const PairArgs& a, cudaStream_t st);
int launch_density_mask_wide(const PairArgs& a, cudaStream_t st);   // R = 2, 3
int launch_force_mask_wide(const PairArgs& a, cudaStream_t st);

int launch_box_to_host(const int* d_box, int* h_box_pinned, cudaStream_t st);   // kernel-written read-back (no copy engine)
int launch_word_to_host(const int* d_word, int* h_word_pinned, cudaStream_t st);
int launch_words_to_host(const uint32_t* d_src, uint32_t* h_dst_pinned, unsigned nwords, cudaStream_t st);
int launch_cfl_probe(DeviceScalars* sc, IntegrateConsts ic, cudaStream_t st);   // sphb_cfl_timestep: sc->dt = the CFL rule, nothing consumed
int launch_integrate(size_t n, float4* posm, float4* velid, const float4* acc, IntegrateConsts ic, DeviceScalars* sc,
                     int* d_box_or_null, cudaStream_t st);
int launch_max_speed(size_t n, const float4* velid, DeviceScalars* sc, cudaStream_t st);

int launch_unpermute(size_t n, const float4* posm, const float4* velid, const float2* rho_p, const float4* acc,
                     const uint64_t* refkeys, const uint32_t* nbr_count, float* pos3, float* vel3, float* rho, float* P,
                     float* acc3, uint64_t* keys, uint32_t* perm, uint32_t* counts, cudaStream_t st);
int launch_export_instances(size_t n, const float4* posm, const float4* velid, const float* colors, const float default_color[3],
                            float* out9, cudaStream_t st);
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
