/* sphb.h — C ABI of the B200-native SPH hot path (libsphb.so).
 *
 * This is the drop-in boundary for ONE path of lucien-vallois/sph-particle-simulator: the body of
 * sph::SPHEngine::step(dt) (reference src/sph_engine.cpp:93-144) and the state accessors around it.
 * The reference has no FFI seam of its own — callers link the concrete C++ class — so the seam is
 * introduced here: a host-side SPHEngine shell (sph-particle-simulator_b200/host/) keeps the
 * reference's public class surface and forwards the hot path through these entry points.
 * Every function cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C, no C++/torch types; all pointers are caller-owned HOST memory unless named d_*
 *     (pinned host memory makes the copies asynchronous DMA, pageable memory works too)
 *   - return 0 on success, a negative SPHB_E_* code on failure; never throws, never aborts;
 *     sphb_last_error() gives the message of the last failure on that context
 *   - a context is single-threaded (like the reference engine) and bound to one CUDA device
 *   - particle order at this interface is ALWAYS the caller's insertion order (reference id = index,
 *     src/particle.cpp:26-33); the device keeps particles cell-sorted internally
 *   - there is NO CPU fallback: without a usable CUDA device sphb_create fails with SPHB_E_CUDA
 */
#ifndef SPHB_H_
#define SPHB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB_VERSION 100

typedef struct sphb_ctx sphb_ctx;

enum {
    SPHB_OK = 0,
    SPHB_E_INVALID = -1,   /* bad argument / call order */
    SPHB_E_CUDA = -2,      /* CUDA runtime failure (message has the CUDA error string) */
    SPHB_E_CAPACITY = -3,  /* more particles than the context capacity */
    SPHB_E_GRID = -4,      /* cell grid implied by bounds / positions exceeds the supported size */
    SPHB_E_NOMEM = -5
};

/* POD mirror of sph::SPHParameters (reference src/sph_engine.h:14-34), same field order and
 * defaults; 16 floats.  neighbor_search_radius is also the hash cell size (sph_engine.cpp:22). */
typedef struct sphb_params {
    float rest_density;           /* 1000   */
    float gas_constant;           /* 2000   */
    float viscosity;              /* 0.001  */
    float smoothing_length;       /* 0.02   */
    float particle_mass;          /* 0.001  */
    float timestep;               /* 0.001  */
    float gravity;                /* -9.81  */
    float damping;                /* 0.99   */
    float CFL_factor;             /* 0.4    */
    float xmin, xmax, ymin, ymax, zmin, zmax; /* -1..1 each */
    float neighbor_search_radius; /* 0.04   */
} sphb_params;

/* Mirror of SPHEngine::PerformanceStats (reference src/sph_engine.h:51-59).  Stage times are GPU
 * times from CUDA events and are only accumulated while SPHB_OPT_STAGE_TIMING is 1;
 * max_neighbors is the running maximum list length including self (sph_engine.cpp:351),
 * total_neighbor_queries grows by N per step (the race-free value of sph_engine.cpp:350). */
typedef struct sphb_stats {
    double total_time;
    double neighbor_search_time;
    double density_computation_time;
    double force_computation_time;
    double integration_time;
    uint64_t max_neighbors;
    uint64_t total_neighbor_queries;
    uint64_t steps;
    uint64_t kernel_launches;      /* kernels of this library launched since the last reset */
    uint64_t error_flags;          /* sticky since the last reset.  SPHB_ERR_OUTSIDE_CELL_BOX (bit 0): a position fell outside
                                    * the internal cell table (NaN/inf, or outside the slab's ghost range) and was binned
                                    * into an edge cell — neighbour sets stay exact (the radius test is), the step is slower */
} sphb_stats;
#define SPHB_ERR_OUTSIDE_CELL_BOX 1u

enum {
    /* 0 = strict: every fp32 operation of density/force/integration is issued in the reference's
     *     association order with IEEE round-to-nearest and no FMA contraction (bit-exact against
     *     the reference compiled without -ffast-math);
     * 1 = fast (default): FMA contraction, reciprocal/rsqrt forms — same neighbour sets, results
     *     within the tolerances stated in DESIGN.md (the reference itself ships with -ffast-math). */
    SPHB_OPT_MATH_MODE = 1,
    /* neighbour-cell walk radius: 1 = 27 cells (default), 2 = 125 cells like SpatialHash::query_squared
     * (reference src/spatial_hash.cpp:35) */
    SPHB_OPT_WALK_RADIUS = 2,
    /* 1 = record per-stage CUDA-event times into sphb_stats (adds event records; default 0) */
    SPHB_OPT_STAGE_TIMING = 3,
    /* 1 = keep per-particle cell keys and neighbour counts of every step for sphb_debug_dump (default 0) */
    SPHB_OPT_DEBUG_CAPTURE = 4,
    /* pair kernel variant (fast mode; strict always runs 0): 2 = the density pass hands the accepted-neighbour
     * sets to the force pass as per-column bitmasks, so the radius test runs once per step (default);
     * 0 = tested per-thread walk in both passes.  Both find identical neighbour sets. */
    SPHB_OPT_PAIR_KERNEL = 5,
    /* fast mode only: the device sorts on an internal grid of cell size neighbor_search_radius / f and walks
     * (2 f + 1)^3 cells, which cuts the candidates per particle (27 r^3 at f = 1, 15.6 r^3 at 2, 11.4 r^3 at 4) and, with
     * ~1 particle per cell, makes the work of the lanes of a warp uniform.  The reference's 63-bit keys, its
     * permutation and the neighbour sets are unaffected (1..6, default 4 — the optimum for h = 2 dx; denser kernels, e.g.
     * h = 2.5 dx, gain from 5: the target is ~1 particle per internal cell; a cell table that would be too large falls
     * back to coarser grids; strict mode always uses 1 so that its layout and summation order are the reference's). */
    SPHB_OPT_GRID_REFINE = 6,
    /* fast mode, pair kernel 2: physical axis (0 = x, 1 = y, 2 = z) that is most significant in the device's
     * cell order (default 0).  In slab mode the slab axis is used, so that ghost layers are contiguous in the
     * sorted arrays.  Only the fp32 summation ORDER depends on it (results agree within the fast-mode
     * tolerances; identical layouts give bit-identical results). */
    SPHB_OPT_LAYOUT_MAJOR = 7,
    /* fast mode, pair kernel 2, internal walk radius >= 4 (grid refine >= 4): where the pair kernels read their
     * candidates from.  0 = every lane loads its candidates from global memory (pair_mask.cu); 1 = staged
     * (pair_stage.cu): every warp brings the neighbour-cell particles of its 32 consecutive cell-sorted particles into
     * shared memory with asynchronous copies, one cell-column group ahead, and the lanes traverse them from there;
     * 2 (default) = the density pass staged, the force pass per lane (the faster combination as measured).
     * Identical results in all three, bit for bit.  Environment override of the default: SPHB_PAIR_MODE. */
    SPHB_OPT_PAIR_MODE = 8,
    /* the smoothing kernel, in the order of the reference's KernelType (src/kernels.h:94-98): 0 = cubic spline (default —
     * what SPHEngine constructs, sph_engine.cpp:17, 23, 162), 1 = Wendland C2, 2 = Gaussian: the other two classes behind
     * create_kernel (src/kernels.cpp:166-236), which the reference engine could host unchanged because it calls
     * W / gradW / laplacianW through a Kernel pointer (sph_engine.cpp:209, 232, 236).  1 and 2 run the tested-walk kernels
     * (the bitmask hand-off kernels are specialised for the cubic spline).  Strict mode: Wendland C2 is bit-exact against
     * the reference classes; the Gaussian agrees to the last bits of expf.  Unlike the reference's engine slot, the choice
     * survives sphb_set_params. */
    SPHB_OPT_KERNEL_TYPE = 9,
    /* 1 (default): steps whose launch sequence repeats exactly (same particle count, parameters, dt argument, grid and
     * buffers; no stage timing, no debug capture) are replayed from a CUDA graph — one launch call per step instead of
     * nine, which is what the step time of small scenes (the reference's own 1 000 - 10 000 particle benchmarks) consists
     * of.  Results are identical: the graph holds the very same kernels.  0 = always launch directly. */
    SPHB_OPT_STEP_GRAPHS = 10,
    /* fast mode, pair kernel 2, internal walk radius >= 4: lanes that share one particle in the density and force passes
     * (1 = default, 2, 4, 8).  With one thread per particle a SMALL scene (the reference's own drivers run 1 000 - 20 000
     * particles) is bound by the serial neighbour walk of a single thread while most of the GPU idles; with L lanes the
     * column groups of a particle are dealt to L lanes and the partial sums meet in shuffle reductions (pair_split.cu).
     * Same neighbour sets and per-pair arithmetic; the per-particle sums are associated differently (fast-mode
     * tolerances).  The host shell selects 8 for particle sets of up to 16 384 and 4 up to 131 072. */
    SPHB_OPT_LANES_PER_PARTICLE = 11
};

/* ---- lifetime ---------------------------------------------------------------------------------
 * replaces SPHEngine::SPHEngine(size_t max_particles) / ~SPHEngine (reference sph_engine.h:91-92,
 * sph_engine.cpp:13-18).  device = CUDA ordinal. */
int sphb_create(sphb_ctx** out, size_t capacity, int device);
void sphb_destroy(sphb_ctx* ctx);
const char* sphb_last_error(const sphb_ctx* ctx); /* ctx may be NULL: error of the last failed sphb_create */
int sphb_version(void);

int sphb_set_option(sphb_ctx* ctx, int option, int64_t value);
int sphb_get_option(const sphb_ctx* ctx, int option, int64_t* value);
/* Run all work of this context on the given cudaStream_t (0/NULL = the legacy default stream). */
int sphb_set_stream(sphb_ctx* ctx, void* cuda_stream);
int sphb_synchronize(sphb_ctx* ctx);

/* ---- configuration ----------------------------------------------------------------------------
 * replaces the parameter state that SPHEngine::initialize / set_parameters / set_gravity /
 * set_viscosity / set_smoothing_length / set_boundaries establish (reference sph_engine.cpp:20-33,
 * 152-172).  The Q1 quirk (cell size = neighbor_search_radius, not 2h) is the caller's to preserve:
 * this call takes the final values verbatim. */
int sphb_set_params(sphb_ctx* ctx, const sphb_params* p);
int sphb_get_params(const sphb_ctx* ctx, sphb_params* p);

/* ---- particle state ---------------------------------------------------------------------------
 * sphb_upload replaces ParticleSystem::clear + add_particles as seen by the hot path (reference
 * particle.cpp:22-39): the device state becomes exactly these n particles, id = index.
 * vel3 == NULL → zero velocities; mass == NULL → params.particle_mass for every particle.
 * The call returns once the work is enqueued.  The host-to-device copies run on a stream of the library's own, so
 * they overlap whatever is still queued on the context's stream (the previous step, a read-back in flight); the
 * kernel that installs the new state is ordered behind both.  Pinned arrays must stay valid until the context's
 * stream has passed that kernel (sphb_synchronize, or any later synchronising call); pageable arrays are consumed
 * before the call returns. */
int sphb_upload(sphb_ctx* ctx, size_t n, const float* pos3, const float* vel3, const float* mass);
/* Strided variant for an array-of-structs such as the reference's 76-byte sph::Particle
 * (particle.h:17-49): byte offsets of position[3], velocity[3], mass inside each record. */
int sphb_upload_strided(sphb_ctx* ctx, size_t n, const void* base, size_t stride,
                        size_t off_pos, size_t off_vel, size_t off_mass);
/* replaces the getters SPHEngine::get_positions/get_velocities/get_densities/get_pressures
 * (sph_engine.h:133-136) and the engine's accelerations_ buffer (sph_engine.h:64).  Each output is
 * in insertion order, n entries (n*3 for vectors); any pointer may be NULL.  Density, pressure and
 * acceleration belong to the last step: between an upload and the next step they read as zeros (the
 * reference would show its previous per-id buffers there). */
int sphb_download(sphb_ctx* ctx, float* pos3, float* vel3, float* rho, float* pressure, float* acc3);
/* The same download in two halves, for callers that stream: _begin un-permutes on the context's stream and starts the
 * copies on a stream of the library's own, then returns; _end waits for them (a second _begin, or sphb_destroy, waits as
 * well).  Whatever the caller enqueues in between — typically the NEXT sphb_upload, whose host-to-device copy runs on the
 * other direction of the PCIe link — overlaps the transfer.  The arrays must stay valid (and should be pinned: pageable
 * memory makes the copies synchronous) until _end.  No reference counterpart (its getters return host copies). */
int sphb_download_begin(sphb_ctx* ctx, float* pos3, float* vel3, float* rho, float* pressure, float* acc3);
int sphb_download_end(sphb_ctx* ctx);
/* Write position/velocity/density/pressure back into an array-of-structs (byte offsets; pass
 * (size_t)-1 for a field to skip). */
int sphb_download_strided(sphb_ctx* ctx, void* base, size_t stride, size_t off_pos, size_t off_vel,
                          size_t off_density, size_t off_pressure);
int sphb_size(const sphb_ctx* ctx, size_t* n);

/* ---- the hot path -----------------------------------------------------------------------------
 * replaces SPHEngine::step(float dt) (reference sph_engine.cpp:93-144): neighbour build, density +
 * EOS, pressure + viscosity force, leapfrog, AABB clamp, time += dt.  dt <= 0 selects the adaptive
 * CFL timestep of SPHEngine::compute_cfl_timestep (sph_engine.cpp:312-333), evaluated on the device
 * including its "particle 0 only" force criterion.  Asynchronous: returns after enqueueing. */
int sphb_step(sphb_ctx* ctx, float dt);
/* replaces SPHEngine::run_steps (sph_engine.cpp:146-150): n steps with fixed dt, or adaptive if dt <= 0 */
int sphb_run_steps(sphb_ctx* ctx, size_t n, float dt);

/* replaces get_current_time / get_step_count (sph_engine.h:111-112); synchronises. */
int sphb_get_time(sphb_ctx* ctx, float* current_time, uint64_t* step_count);
/* clear_particles resets time and step count, initialize_* do not (sph_engine.cpp:47,87-91) */
int sphb_set_time(sphb_ctx* ctx, float current_time, uint64_t step_count);
/* the dt the NEXT adaptive step would use (compute_cfl_timestep, sph_engine.cpp:312-333); synchronises */
int sphb_cfl_timestep(sphb_ctx* ctx, float* dt);

/* ---- diagnostics ------------------------------------------------------------------------------
 * replaces get_performance_stats / reset_performance_stats (sph_engine.h:124-125) */
int sphb_get_stats(sphb_ctx* ctx, sphb_stats* out);
int sphb_reset_stats(sphb_ctx* ctx);
/* Device reductions behind get_total_mass / get_total_energy / compute_conservation_errors
 * (sph_engine.cpp:178-200), accumulated in fp64: sum_density = Σ rho_i (caller multiplies by h^3),
 * kinetic = Σ 0.5 m_i |v_i|^2, max_speed = max |v_i|.  Any pointer may be NULL. */
int sphb_diagnostics(sphb_ctx* ctx, double* sum_density, double* kinetic, float* max_speed);

/* replaces the per-frame array-of-structs walk of Renderer::update_particle_data (reference src/renderer.cpp:279-312):
 * writes the renderer's instance records — position (3), velocity (3), colour (3) floats per particle, insertion
 * order, n * 9 floats — with one kernel.  dst_on_device != 0: dst is a DEVICE pointer (e.g. a CUDA-mapped vertex
 * buffer) and nothing crosses the bus; otherwise dst is host memory.  Colours: sphb_set_colors uploads per-particle RGB
 * once (sph::Particle::color is never changed by the physics); without it every record carries default_rgb (NULL:
 * the reference's Particle default 0, 0.5, 1). */
int sphb_set_colors(sphb_ctx* ctx, size_t n, const float* rgb3);
int sphb_export_instances(sphb_ctx* ctx, float* dst, int dst_on_device, const float* default_rgb);

/* Parity hooks (need SPHB_OPT_DEBUG_CAPTURE = 1 before the step), all for the LAST step, i.e. for
 * the positions that step's neighbour build saw:
 *   keys[i]      63-bit cell key of particle i — SpatialHash::hash_position (spatial_hash.h:20-36)
 *   perm[s]      id of the particle in sorted slot s == std::stable_sort of ids by keys[]
 *   nbr_count[i] neighbour-list length of particle i, self included (spatial_hash.cpp:31-57)
 * Any pointer may be NULL. */
int sphb_debug_dump(sphb_ctx* ctx, uint64_t* keys, uint32_t* perm, uint32_t* nbr_count);
/* Test hook, no device needed: the static spherical stencil of the fast path for walk radius `radius` (2..6 = walk
 * radius x grid refine).  reach[(d0 + radius) * (2 radius + 1) + d1 + radius] = largest |d2| visited in cell column
 * (d0, d1), -1 = column skipped; *cell_scale = factor applied to refine / neighbor_search_radius to get the internal
 * 1 / cell size.  Returns the number of columns, or SPHB_E_INVALID.  (Replaces nothing in the reference: its
 * SpatialHash::query_squared, spatial_hash.cpp:31-57, visits the full cube of cells.) */
int sphb_debug_stencil(int radius, int8_t* reach, float* cell_scale);

/* ---- slab decomposition across GPUs (one context per GPU) ------------------------------------------
 * NEW, no reference counterpart: the reference is a single-process CPU program.  The domain is cut
 * into slabs of whole reference cells along one axis; each context owns [own_lo, own_hi) and, before a
 * step, also holds copies ("ghosts") of the halo_layers cell layers beyond each face.  With two layers
 * a step needs ONE halo exchange: densities of the first layer are recomputed locally from the second.
 * The host driver (sph-particle-simulator_b200/slab.py) moves the 32-byte records
 * {x, y, z, mass, vx, vy, vz, id} between contexts over NCCL (torch.distributed); d_* pointers below are
 * DEVICE memory owned by the caller.  Because every step re-sorts owned + ghost particles by
 * (cell, global id), results do not depend on the number of slabs. */
typedef struct sphb_slab {
    int32_t axis;                   /* 0 = x, 1 = y, 2 = z */
    int32_t own_lo, own_hi;         /* owned reference cells [own_lo, own_hi) on that axis */
    int32_t halo_layers;            /* ghost layers per face (>= 2) */
    uint64_t id_space;              /* global ids are < id_space (< 2^31) */
    float box_min[3], box_max[3];   /* global box containing every position that can occur */
} sphb_slab;

int sphb_set_slab(sphb_ctx* ctx, const sphb_slab* slab);   /* NULL: back to the whole-domain mode */
/* like sphb_upload, with explicit global ids (slab mode: the particles this context owns) */
int sphb_upload_ids(sphb_ctx* ctx, size_t n, const float* pos3, const float* vel3, const float* mass, const uint32_t* ids);
/* One-round exchange: drop all ghosts, then route every owned particle to the rank that
 * owns its cell now AND, flagged as ghost (id bit 31), to each adjacent rank whose halo layers contain the
 * cell.  d_out receives, for r = 0..nranks-1, [records owned by r (none for r = my_rank)][ghosts for r];
 * counts[2r] / counts[2r+1] are those group sizes (counts[2*my_rank] = particles kept in place).
 * Append what arrives with sphb_slab_append(..., ghost = -1): records keep their own flag.  Synchronises. */
int sphb_slab_exchange_pack(sphb_ctx* ctx, const int32_t* cuts, int nranks, int my_rank, void* d_out, size_t cap_records,
                            uint64_t* counts);
/* The same exchange in two phases, for ONE host synchronisation per step: _count only enqueues the routing count
 * and copies this rank's 2*nranks group sizes (uint32) to the DEVICE buffer d_counts on the context's stream — the
 * caller all-gathers them on the device and reads the gathered table back once; _split then takes this rank's row
 * of that table from the host and moves the records exactly like sphb_slab_exchange_pack.  Nothing may touch the
 * context between the two calls. */
int sphb_slab_exchange_count(sphb_ctx* ctx, const int32_t* cuts, int nranks, int my_rank, uint32_t* d_counts);
int sphb_slab_exchange_split(sphb_ctx* ctx, const int32_t* cuts, int nranks, int my_rank, const uint32_t* counts, void* d_out,
                             size_t cap_records);
/* Append records: ghost = -1 keeps each record's own flag (what the exchange delivers); 0 / 1 force owned / ghost. */
int sphb_slab_append(sphb_ctx* ctx, const void* d_in, size_t count, int ghost);
/* Helper of the exchange driver: a small DEVICE buffer (the gathered table of group sizes) into PINNED host memory, written
 * by a kernel on the context's stream instead of a device-to-host copy — a copy would queue on the copy engine behind a
 * bulk read-back that is still in flight (sphb_slab_download_begin) and hold the step up behind it.  Enqueue only: the
 * caller synchronises the stream before reading.  bytes: a multiple of 4, at most 1 MiB. */
int sphb_read_small(sphb_ctx* ctx, const void* d_src, void* h_dst_pinned, size_t bytes);
/* Owned particles of this context in arbitrary order: ids[k] with the matching fields (host pointers,
 * any field may be NULL); *count = number written (<= cap). */
/* Adaptive timestep across slabs: compute_cfl_timestep (reference sph_engine.cpp:312-333) needs the global
 * max |v|^2 and the acceleration of particle id 0.  get: this context's max |v|^2 over its owned particles, its
 * copy of a0 and whether it advanced particle 0 in the last step (a0_fresh; cleared by the call).  set: install
 * the globally reduced values before an adaptive sphb_step.  Both synchronise. */
int sphb_get_cfl_state(sphb_ctx* ctx, float* max_v2, float* a0_xyz, int* a0_fresh);
int sphb_set_cfl_state(sphb_ctx* ctx, float max_v2, const float* a0_xyz);
int sphb_slab_download(sphb_ctx* ctx, size_t cap, uint32_t* ids, float* pos3, float* vel3, float* rho, float* pressure,
                       float* acc3, size_t* count);
/* The same in two halves (see sphb_download_begin): _begin enqueues the export and the copies — min(cap, particles held
 * incl. halo copies) entries per field, because the number of OWNED particles is only known on the device at that point —
 * and returns; _end waits and reports how many leading entries are owned particles. */
int sphb_slab_download_begin(sphb_ctx* ctx, size_t cap, uint32_t* ids, float* pos3, float* vel3, float* rho, float* pressure,
                             float* acc3);
int sphb_slab_download_end(sphb_ctx* ctx, size_t* count);

/* ---- several GPUs of one node behind ONE handle (csrc/multi.cu) -----------------------------------------
 * "Multi-GPU: sphb_create_multi(ctx**, capacity, ndev, const int* devs) with the same calls".
 * NEW, no reference counterpart (the reference is a single-process CPU program): one host thread drives one
 * context per device with the slab protocol above — cuts of whole reference cells along one axis balanced by
 * particle count at upload time, ONE exchange round per step (migration + two halo layers), records moved by
 * peer-to-peer copies over NVLink (cudaMemcpyPeerAsync), no NCCL and no second process.  Every call mirrors
 * the single-context call of the same name and keeps its contract: insertion order at the interface, host
 * pointers, 0 / negative SPHB_E_* codes.  Strict mode reproduces the single-context results bit for bit for
 * any number of devices; fast mode does when the single context is given the same SPHB_OPT_LAYOUT_MAJOR.
 * `devices` may name the same ordinal more than once (several slabs on one GPU: how the path is tested on
 * a one-GPU box).  ndev = 1 is the plain single-context path.  This is what lets sph::SPHEngine
 * (reference sph_engine.h:89-141) own N GPUs: the host shell selects it with SPHB_DEVICES=0,1,... */
typedef struct sphb_multi sphb_multi;
enum {
    /* slab axis: -1 (default) = the longest axis of the uploaded particles' bounding box, else 0 / 1 / 2 */
    SPHB_OPT_MULTI_AXIS = 100,
    /* ghost layers per slab face (default 2 = the minimum: the first layer's density is recomputed locally from the second);
     * takes effect at the next upload */
    SPHB_OPT_MULTI_HALO_LAYERS = 101,
    /* load balance: when an exchange leaves one device with more than 1.5x its even share AND more than this many
     * particles above it (default 32768), the slabs are re-cut for the current positions before the next step (one
     * round trip of positions and velocities through the host; results do not depend on where the cuts are) */
    SPHB_OPT_MULTI_REBALANCE_MIN = 102
};
int sphb_create_multi(sphb_multi** out, size_t capacity, int ndev, const int* devices);
void sphb_destroy_multi(sphb_multi* m);
const char* sphb_multi_last_error(const sphb_multi* m);   /* m may be NULL: error of the last failed sphb_create_multi */
int sphb_multi_device_count(const sphb_multi* m);
/* any SPHB_OPT_* of the single context (applied to every device) or SPHB_OPT_MULTI_* */
int sphb_multi_set_option(sphb_multi* m, int option, int64_t value);
int sphb_multi_set_params(sphb_multi* m, const sphb_params* p);
/* needs the parameters first: the slabs are cut in units of neighbor_search_radius */
int sphb_multi_upload(sphb_multi* m, size_t n, const float* pos3, const float* vel3, const float* mass);
int sphb_multi_upload_strided(sphb_multi* m, size_t n, const void* base, size_t stride, size_t off_pos, size_t off_vel,
                              size_t off_mass);
int sphb_multi_step(sphb_multi* m, float dt);             /* dt <= 0: the adaptive timestep from the global maxima */
int sphb_multi_run_steps(sphb_multi* m, size_t n, float dt);
int sphb_multi_synchronize(sphb_multi* m);
int sphb_multi_size(sphb_multi* m, size_t* n);
int sphb_multi_download(sphb_multi* m, float* pos3, float* vel3, float* rho, float* pressure, float* acc3);
int sphb_multi_download_strided(sphb_multi* m, void* base, size_t stride, size_t off_pos, size_t off_vel, size_t off_density,
                                size_t off_pressure);
int sphb_multi_get_time(sphb_multi* m, float* current_time, uint64_t* step_count);
int sphb_multi_set_time(sphb_multi* m, float current_time, uint64_t step_count);
int sphb_multi_cfl_timestep(sphb_multi* m, float* dt);
/* stage times: the maximum over the devices; kernel_launches: the sum; max_neighbors: the maximum */
int sphb_multi_get_stats(sphb_multi* m, sphb_stats* out);
int sphb_multi_reset_stats(sphb_multi* m);
int sphb_multi_diagnostics(sphb_multi* m, double* sum_density, double* kinetic, float* max_speed);
/* per device: particles owned and ghost copies held after the last step's exchange (arrays of ndev, may be NULL) */
int sphb_multi_layout(sphb_multi* m, int32_t* cuts_ndev_plus_1, int* axis, uint64_t* owned, uint64_t* ghosts);

#ifdef __cplusplus
}
#endif
#endif /* SPHB_H_ */
