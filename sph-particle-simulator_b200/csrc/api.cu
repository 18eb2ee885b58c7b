// C ABI of libsphb.so (include/sphb.h): context, buffers, step orchestration.
// Host-side C++ only touches plain pointers and CUDA runtime calls; there is no CPU compute path.
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "sphb_internal.cuh"

namespace sphb {
int launch_unpack_strided(size_t n, const unsigned char* d_base, size_t stride, size_t off_pos, size_t off_vel, size_t off_mass,
                          float4* posm, float4* velid, DeviceScalars* sc, cudaStream_t st);
}

using namespace sphb;

namespace {
thread_local char g_create_error[512] = "";
constexpr uint64_t kMaxCells = 1ull << 30;   // dense cell table of up to 4 GB (a 100 M-particle scene on one GPU needs 3.1e8 internal cells)
constexpr int kMaxCoord = 1 << 20;  // |cell coordinate| limit for an injective reference key
}  // namespace

struct sphb_ctx {
    int device = 0;
    size_t capacity = 0;
    size_t n = 0;
    cudaStream_t stream = nullptr;

    sphb_params prm{};
    bool have_params = false;

    int math_mode = 1;
    int walk_radius = 1;
    int stage_timing = 0;
    int debug_capture = 0;
    int pair_kernel = 2;     // fast mode: 2 = bitmask hand-off density -> force (default), 0 = tested walk twice
    int pair_mode = 2;       // R >= 4 mask kernels: 0 = per-lane global loads (pair_mask.cu), 1 = shared-memory staged (pair_stage.cu), 2 = staged density pass + per-lane force pass
    int layout_major = 0;    // fast-mode layout: physical axis that is most significant in the cell order (slab mode: the slab axis)
    int grid_refine = 4;     // internal cell = neighbor_search_radius / grid_refine (fast mode; strict always 1)
    int lanes = 1;           // SPHB_OPT_LANES_PER_PARTICLE: lanes sharing one particle in the bitmask pair kernels (small scenes)
    int kernel_type = 0;     // SPHB_OPT_KERNEL_TYPE: 0 cubic spline (what SPHEngine constructs), 1 Wendland C2, 2 Gaussian

    float4* posm[2] = {nullptr, nullptr};
    float4* velid[2] = {nullptr, nullptr};
    int cur = 0;
    float2* rho_p = nullptr;
    float4* fa = nullptr;    // staging of the tested-walk kernels (strict mode, variant 0): allocated on first use
    float4* fb = nullptr;
    float4* acc = nullptr;
    void* masks = nullptr;       // variant 2: (mask_cols(R) + 1) rows of mask_stride accepted-neighbour masks
    size_t mask_bytes = 0;
    size_t mask_stride = 0;
    ForceRec* fab = nullptr;     // variant 2: 32-byte force-pass records
    float* colors = nullptr;     // optional per-id RGB of the renderer's instance records (sphb_set_colors)
    size_t n_colors = 0;
    uint32_t* nbr_count = nullptr;
    uint64_t* refkeys[2] = {nullptr, nullptr};
    uint64_t* dbg_keys[2] = {nullptr, nullptr};   // reference-order composite keys (debug capture with a refined grid)
    int dbg_sorted = -1;                          // which dbg_keys buffer holds the sorted keys of the last step, -1: layout order is the reference order
    int dbg_id_bits = 0;
    uint2* cell_ticket = nullptr;   // counting sort: {cell, ticket inside the cell} per particle
    uint32_t* slot_src = nullptr;   // counting sort: particle that landed in each (cell-ordered) slot
    uint32_t* dbg_vals[2] = {nullptr, nullptr};   // payload buffers of the debug radix sort
    uint32_t* cell_start = nullptr; // dense cell table (ncells + 1 entries) followed by the scratch of its scan
    size_t cell_cap = 0;            // bytes
    DeviceScalars* sc = nullptr;
    DeviceScalars* h_sc = nullptr;  // pinned mirror for read-back
    unsigned char* d_stage = nullptr;
    size_t stage_bytes = 0;
    // sphb_upload: the host-to-device copies run on a stream of their own into a staging buffer of their own, so that they
    // overlap whatever is still running on the context's stream (typically the previous step); only the pack kernel waits
    float* d_in = nullptr;
    size_t in_bytes = 0;
    cudaStream_t in_stream = nullptr;
    cudaEvent_t ev_in_done = nullptr, ev_in_free = nullptr;
    bool in_used = false;
    // sphb_download_begin / _end: un-permuted fields staged here and copied out on a stream of their own, so that the
    // copy overlaps whatever the caller enqueues next (typically the next upload: PCIe is full duplex)
    float* d_out = nullptr;
    size_t out_bytes = 0;
    cudaStream_t out_stream = nullptr;
    cudaEvent_t ev_out_ready = nullptr, ev_out_done = nullptr;
    bool out_pending = false;
    size_t out_cap = 0;        // sphb_slab_download_begin: capacity of the caller's arrays
    unsigned char* h_bounce = nullptr;  // pinned bounce buffer for the strided download
    size_t bounce_bytes = 0;

    // float bounding box of the positions currently on the device (valid when n > 0)
    float box_min[3] = {0, 0, 0}, box_max[3] = {0, 0, 0};
    bool box_pending = false;  // box must be read back from the device (after an upload)
    int* d_box = nullptr;      // 6 ordered-int encoded floats
    int* h_box = nullptr;      // pinned host copy, written by a kernel (fetch_box)
    bool stepped_since_upload = false;
    bool box_tracking = false;   // the last step's integrate kernel reduced the particle bounding box into d_box

    bool slab_on = false;
    sphb_slab slab{};
    void* sort_scratch = nullptr;       // debug radix sort: global histograms, look-back status words, tickets
    unsigned int* d_counts = nullptr;   // 2 * kMaxRanks + 1 counters / cursors

    uint64_t step_count = 0;
    sphb_stats stats{};
    // stage-timing events: one set of 5 per step, resolved lazily (no sync inside sphb_step)
    struct EventSet { cudaEvent_t e[5]; };
    std::vector<EventSet> ev_pool;
    size_t ev_used = 0;

    // CUDA graphs of repeating steps (sphb_step), one per parity of the double-buffered arrays
    struct StepGraph {
        cudaGraphExec_t exec = nullptr;
        std::vector<unsigned char> key, seen;   // configuration the graph was captured for / seen at the last eligible step
        uint64_t launches = 0;
    };
    StepGraph graphs[2];
    cudaStream_t capture_stream = nullptr;      // graphs are captured here and launched into `stream`
    int use_graphs = 1;                         // SPHB_OPT_STEP_GRAPHS

    char err[512] = "";
};

namespace {

// bytes of everything a step's launches depend on (compared to decide whether a captured graph can be replayed)
struct StepKey {
    std::vector<unsigned char> bytes;
    template <typename T> void add(const T& v) {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(&v);
        bytes.insert(bytes.end(), p, p + sizeof(T));
    }
};

int fail(sphb_ctx* c, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    if (c) vsnprintf(c->err, sizeof(c->err), fmt, ap);
    else vsnprintf(g_create_error, sizeof(g_create_error), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(c, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail((c), SPHB_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int bits_for(uint64_t count) {  // bits needed to represent values 0..count-1
    int b = 1;
    while (b < 63 && (1ull << b) < count) ++b;
    return b;
}

int ensure_stage(sphb_ctx* c, size_t bytes) {
    if (bytes <= c->stage_bytes) return SPHB_OK;
    if (c->d_stage) cudaFree(c->d_stage);
    c->d_stage = nullptr;
    c->stage_bytes = 0;
    CU(c, cudaMalloc(&c->d_stage, bytes));
    c->stage_bytes = bytes;
    return SPHB_OK;
}

int ensure_bounce(sphb_ctx* c, size_t bytes) {
    if (bytes <= c->bounce_bytes) return SPHB_OK;
    if (c->h_bounce) cudaFreeHost(c->h_bounce);
    c->h_bounce = nullptr;
    c->bounce_bytes = 0;
    CU(c, cudaMallocHost(&c->h_bounce, bytes));
    c->bounce_bytes = bytes;
    return SPHB_OK;
}

int ensure_debug(sphb_ctx* c) {
    if (c->refkeys[0]) return SPHB_OK;
    const size_t cap = c->capacity ? c->capacity : 1;
    CU(c, cudaMalloc(&c->refkeys[0], cap * sizeof(uint64_t)));
    CU(c, cudaMalloc(&c->refkeys[1], cap * sizeof(uint64_t)));
    CU(c, cudaMalloc(&c->nbr_count, cap * sizeof(uint32_t)));
    CU(c, cudaMalloc(&c->dbg_keys[0], cap * sizeof(uint64_t)));
    CU(c, cudaMalloc(&c->dbg_keys[1], cap * sizeof(uint64_t)));
    CU(c, cudaMalloc(&c->dbg_vals[0], cap * sizeof(uint32_t)));
    CU(c, cudaMalloc(&c->dbg_vals[1], cap * sizeof(uint32_t)));
    CU(c, cudaMalloc(&c->sort_scratch, onesweep_scratch_bytes(cap, 64)));
    CU(c, cudaMemset(c->refkeys[0], 0, cap * sizeof(uint64_t)));
    CU(c, cudaMemset(c->refkeys[1], 0, cap * sizeof(uint64_t)));
    CU(c, cudaMemset(c->nbr_count, 0, cap * sizeof(uint32_t)));
    return SPHB_OK;
}

// ordered-int decoding of the device bounding box (see k_pack_upload's encode)
float decode_ordered(int v) {
    int s = v >= 0 ? v : v ^ 0x7FFFFFFF;
    float f;
    memcpy(&f, &s, 4);
    return f;
}

int fetch_box(sphb_ctx* c) {
    if (!c->box_pending) return SPHB_OK;
    // written into pinned host memory by a kernel, not by a device-to-host copy: a copy would queue on the copy engine
    // behind a read-back that is still in flight on another stream (sphb_download_begin) and stall the step behind it
    if (!c->h_box) CU(c, cudaMallocHost(&c->h_box, 8 * sizeof(int)));
    c->stats.kernel_launches += sphb::launch_box_to_host(c->d_box, c->h_box, c->stream);
    CU(c, cudaStreamSynchronize(c->stream));
    const int* h = c->h_box;
    for (int a = 0; a < 3; ++a) {
        c->box_min[a] = decode_ordered(h[a]);
        c->box_max[a] = decode_ordered(h[3 + a]);
    }
    c->box_pending = false;
    return SPHB_OK;
}

int host_cell(float p, float inv_cell) {
    float f = floorf(p * inv_cell);
    if (!(f == f)) return 0;
    if (f < -2147483000.0f) return -2147483647;
    if (f > 2147483000.0f) return 2147483647;
    return (int)f;
}

// layout_major < 0: the reference layout (x, y, z key order, masked-key ranks).  layout_major = a: fast-mode layout
// with physical axis a most significant and monotone ranks (see GridDesc).
int make_grid(sphb_ctx* c, GridDesc* g, int refine, int layout_major = -1, int pad = 0) {
    const float cell = c->prm.neighbor_search_radius;
    if (!(cell > 0.0f)) return fail(c, SPHB_E_INVALID, "neighbor_search_radius must be > 0 (got %g)", (double)cell);
    g->ref_inv_cell = 1.0f / cell;  // SpatialHash::set_cell_size, reference spatial_hash.h:63-66
    g->inv_cell = g->ref_inv_cell * (float)refine;
    // Refined (internal) grids use cells 0.1 % LARGER than neighbor_search_radius / refine: two particles exactly
    // neighbor_search_radius apart (the q = 2 ties of lattice scenes) then differ by strictly less than `refine`
    // in p * inv_cell, so fp32 rounding of that product can never put them refine + 1 cells apart and outside the
    // (2 refine + 1)-cell walk.  The reference guards the same tie with a second ring of cells (spatial_hash.cpp:35).
    if (refine > 1) g->inv_cell *= kRefinedCellScale;
    uint64_t ncells = 1;
    for (int a = 0; a < 3; ++a) {
        int lo = host_cell(c->box_min[a], g->inv_cell), hi = host_cell(c->box_max[a], g->inv_cell);
        if (hi < lo) { int t = lo; lo = hi; hi = t; }
        if (c->slab_on && a == c->slab.axis) {
            // only the owned reference cells and their ghost layers can be populated: the internal cells of the positions
            // at their two faces, by the same formula as the keys (internal cells are 0.1 % larger than
            // neighbor_search_radius / refine), +-1 cell of slack for the roundings of the two products
            const int slo = host_cell((float)(c->slab.own_lo - c->slab.halo_layers) * cell, g->inv_cell) - 1;
            const int shi = host_cell((float)(c->slab.own_hi + c->slab.halo_layers) * cell, g->inv_cell) + 1;
            if (lo < slo) lo = slo;
            if (hi > shi) hi = shi;
            if (hi < lo) hi = lo;
        }
        if (lo < -kMaxCoord || hi >= kMaxCoord)
            return fail(c, SPHB_E_GRID, "cell coordinates [%d, %d] on axis %d exceed the 21-bit key range", lo, hi, a);
        // the 0.1 % cell margin of a refined grid must dominate the fp32 rounding of p * inv_cell (relative 2^-24 of the
        // coordinate): beyond 2^14 cells from the origin fall back to a coarser grid (the caller retries with refine - 1)
        if (refine > 1 && (lo < -(1 << 14) || hi > (1 << 14)))
            return fail(c, SPHB_E_GRID, "refined grid: cell coordinates [%d, %d] too far from the origin for the tie margin", lo, hi);
        lo -= pad;   // guard cells of the fast-mode layout (always empty: positions are clamped into the box)
        hi += pad;
        g->lo[a] = lo;
        g->hi[a] = hi;
        g->ext[a] = hi - lo + 1;
        g->pos_lo[a] = lo > 0 ? lo : 0;
        g->npos[a] = hi >= g->pos_lo[a] ? hi - g->pos_lo[a] + 1 : 0;
        ncells *= (uint64_t)g->ext[a];
        if (ncells > kMaxCells)
            return fail(c, SPHB_E_GRID, "dense cell table would need more than %llu cells (box / neighbor_search_radius too large)",
                        (unsigned long long)kMaxCells);
    }
    g->monotone = layout_major >= 0 ? 1 : 0;
    g->pad = pad;
    g->perm[0] = 0; g->perm[1] = 1; g->perm[2] = 2;
    if (layout_major == 1) { g->perm[0] = 1; g->perm[1] = 0; g->perm[2] = 2; }        // (y, x, z)
    else if (layout_major == 2) { g->perm[0] = 2; g->perm[1] = 0; g->perm[2] = 1; }   // (z, x, y)
    if (g->perm[0] != 0) {   // the loop above filled the fields by physical axis: bring them into key order
        GridDesc t = *g;
        for (int k = 0; k < 3; ++k) {
            const int a = g->perm[k];
            g->lo[k] = t.lo[a]; g->hi[k] = t.hi[a]; g->ext[k] = t.ext[a]; g->pos_lo[k] = t.pos_lo[a]; g->npos[k] = t.npos[a];
        }
    }
    g->ncells = (uint32_t)ncells;
    g->id_bits = bits_for(c->slab_on ? (c->slab.id_space > 1 ? c->slab.id_space : 2) : (c->capacity > 1 ? c->capacity : 2));
    g->cell_bits = bits_for(ncells > 1 ? ncells : 2);
    if (g->id_bits + g->cell_bits > 64) return fail(c, SPHB_E_GRID, "composite sort key exceeds 64 bits");
    return SPHB_OK;
}

// bytes of the cell table: ncells + 1 counters (rounded up to a 16-byte multiple) + the scratch of their scan
size_t cell_table_bytes(const GridDesc& g, size_t* scratch_offset) {
    const size_t entries = ((size_t)g.ncells + 1 + 3) & ~(size_t)3;
    if (scratch_offset) *scratch_offset = entries * sizeof(uint32_t);
    return entries * sizeof(uint32_t) + scan_scratch_bytes((size_t)g.ncells + 1);
}

int ensure_cell_table(sphb_ctx* c, const GridDesc& g) {
    const size_t need = cell_table_bytes(g, nullptr);
    if (need > c->cell_cap) {
        if (c->cell_start) { CU(c, cudaStreamSynchronize(c->stream)); cudaFree(c->cell_start); }
        c->cell_start = nullptr;
        c->cell_cap = 0;
        CU(c, cudaMalloc(&c->cell_start, need));
        c->cell_cap = need;
    }
    return SPHB_OK;
}

PairConsts make_pair_consts(const sphb_params& p, int kernel_type) {
    PairConsts k;
    const float h = p.smoothing_length;
    k.h = h;
    k.h_sq = h * h;                                   // kernels.cpp:13
    k.sigma = 1.0f / (static_cast<float>(M_PI) * h * h * h);  // kernels.cpp:27
    k.r2 = p.neighbor_search_radius * p.neighbor_search_radius;  // sph_engine.cpp:347
    k.w0 = k.sigma * (2.0f / 3.0f);                   // W(0): sigma * (2/3 - 0*0 + 0.5*0*0*0)
    k.wnorm = 21.0f / (2.0f * static_cast<float>(M_PI) * h * h * h);             // kernels.cpp:168
    k.gssi = 1.0f / (h * h);                                                      // kernels.cpp:203
    k.gnorm = 1.0f / powf(static_cast<float>(M_PI) * h * h, 1.5f);                // kernels.cpp:204 (std::pow(float, float))
    // W(0) of the other classes: Wendland norm * 1^4 * (2*0 + 1) (kernels.cpp:171-177), Gaussian norm * exp(-0) (207-210)
    if (kernel_type == kKernelWendlandC2) k.w0 = k.wnorm;
    if (kernel_type == kKernelGaussian) k.w0 = k.gnorm;
    k.rest_density = p.rest_density;
    k.gas_constant = p.gas_constant;
    k.viscosity = p.viscosity;
    k.gravity = p.gravity;
    k.inv_h = 1.0f / h;
    k.sig_h = k.sigma / h;
    k.sig_h2 = k.sigma / k.h_sq;
    k.neg_zero = -0.0f;
    k.r2_next = nextafterf(k.r2, INFINITY);
    return k;
}

IntegrateConsts make_integrate_consts(const sphb_params& p) {
    IntegrateConsts ic;
    ic.damping = p.damping;
    ic.xmin = p.xmin; ic.xmax = p.xmax; ic.ymin = p.ymin; ic.ymax = p.ymax; ic.zmin = p.zmin; ic.zmax = p.zmax;
    ic.cfl = p.CFL_factor;
    ic.h = p.smoothing_length;
    ic.timestep = p.timestep;
    return ic;
}

void free_all(sphb_ctx* c) {
    cudaSetDevice(c->device);
    for (auto& sg : c->graphs) if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
    if (c->capture_stream) { cudaStreamDestroy(c->capture_stream); c->capture_stream = nullptr; }
    if (c->in_stream) { cudaStreamSynchronize(c->in_stream); cudaStreamDestroy(c->in_stream); c->in_stream = nullptr; }
    if (c->ev_in_done) cudaEventDestroy(c->ev_in_done);
    if (c->ev_in_free) cudaEventDestroy(c->ev_in_free);
    cudaFree(c->d_in);
    if (c->out_stream) { cudaStreamSynchronize(c->out_stream); cudaStreamDestroy(c->out_stream); c->out_stream = nullptr; }
    if (c->ev_out_ready) cudaEventDestroy(c->ev_out_ready);
    if (c->ev_out_done) cudaEventDestroy(c->ev_out_done);
    cudaFree(c->d_out);
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->posm[i]); cudaFree(c->velid[i]); cudaFree(c->refkeys[i]); cudaFree(c->dbg_keys[i]); cudaFree(c->dbg_vals[i]);
    }
    cudaFree(c->cell_ticket); cudaFree(c->slot_src);
    cudaFree(c->masks); cudaFree(c->fab); cudaFree(c->colors);
    cudaFree(c->rho_p); cudaFree(c->fa); cudaFree(c->fb); cudaFree(c->acc); cudaFree(c->nbr_count);
    cudaFree(c->cell_start);
    cudaFree(c->sc); cudaFree(c->d_stage); cudaFree(c->d_box); cudaFree(c->d_counts); cudaFree(c->sort_scratch);
    if (c->h_sc) cudaFreeHost(c->h_sc);
    if (c->h_box) cudaFreeHost(c->h_box);
    if (c->h_bounce) cudaFreeHost(c->h_bounce);
    for (auto& set : c->ev_pool) for (auto& e : set.e) if (e) cudaEventDestroy(e);
    c->ev_pool.clear();
}

// Resolve all pending stage-timing event sets into the stats accumulators.
void drain_events(sphb_ctx* c) {
    for (size_t i = 0; i < c->ev_used; ++i) {
        auto& set = c->ev_pool[i];
        if (cudaEventSynchronize(set.e[4]) != cudaSuccess) continue;
        float ms[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&ms[k], set.e[k], set.e[k + 1]);
        c->stats.neighbor_search_time += 1e-3 * ms[0];
        c->stats.density_computation_time += 1e-3 * ms[1];
        c->stats.force_computation_time += 1e-3 * ms[2];
        c->stats.integration_time += 1e-3 * ms[3];
        c->stats.total_time += 1e-3 * (ms[0] + ms[1] + ms[2] + ms[3]);
    }
    c->ev_used = 0;
}

cudaEvent_t* next_event_set(sphb_ctx* c) {
    if (c->ev_used == c->ev_pool.size()) {
        if (c->ev_pool.size() >= 512) {
            drain_events(c);
        } else {
            sphb_ctx::EventSet set{};
            for (auto& e : set.e)
                if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            c->ev_pool.push_back(set);
        }
    }
    return c->ev_pool[c->ev_used++].e;
}

int read_scalars(sphb_ctx* c) {
    CU(c, cudaMemcpyAsync(c->h_sc, c->sc, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return SPHB_OK;
}

}  // namespace

extern "C" {

int sphb_version(void) { return SPHB_VERSION; }

const char* sphb_last_error(const sphb_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int sphb_create(sphb_ctx** out, size_t capacity, int device) {
    if (!out) return fail(nullptr, SPHB_E_INVALID, "sphb_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SPHB_E_CUDA, "no usable CUDA device (%s); libsphb has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fail(nullptr, SPHB_E_INVALID, "device %d out of range [0, %d)", device, ndev);
    if (capacity >= (1ull << 31)) return fail(nullptr, SPHB_E_CAPACITY, "capacity must be below 2^31");
    sphb_ctx* c = new (std::nothrow) sphb_ctx();
    if (!c) return fail(nullptr, SPHB_E_NOMEM, "out of host memory");
    c->device = device;
    c->capacity = capacity;
    if (const char* e = getenv("SPHB_PAIR_MODE")) { if (e[0] >= '0' && e[0] <= '2') c->pair_mode = e[0] - '0'; }   // A/B runs without touching the caller
    if (const char* e = getenv("SPHB_LANES")) { if (e[0] == '1' || e[0] == '2' || e[0] == '4' || e[0] == '8') c->lanes = e[0] - '0'; }
    const size_t cap = capacity ? capacity : 1;
#define CUC(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e2_ = (call);                                                                         \
        if (e2_ != cudaSuccess) {                                                                         \
            fail(nullptr, SPHB_E_CUDA, "%s: %s", #call, cudaGetErrorString(e2_));                         \
            free_all(c);                                                                                  \
            delete c;                                                                                     \
            return SPHB_E_CUDA;                                                                           \
        }                                                                                                 \
    } while (0)
    CUC(cudaSetDevice(device));
    for (int i = 0; i < 2; ++i) {
        // + 4 records: the pair kernels read candidates two at a time, one pair ahead, and may touch up to three
        // records behind the last particle (k_reorder keeps slot n a massless, finite sentinel; the rest stay finite)
        CUC(cudaMalloc(&c->posm[i], (cap + 4) * sizeof(float4)));
        CUC(cudaMemset(c->posm[i], 0, (cap + 4) * sizeof(float4)));
        CUC(cudaMalloc(&c->velid[i], cap * sizeof(float4)));
    }
    CUC(cudaMalloc(&c->cell_ticket, cap * sizeof(uint2)));
    CUC(cudaMalloc(&c->slot_src, cap * sizeof(uint32_t)));
    CUC(cudaMalloc(&c->rho_p, cap * sizeof(float2)));
    CUC(cudaMalloc(&c->acc, cap * sizeof(float4)));
    // densities_/pressures_/accelerations_ are value-initialised by initialize() (sph_engine.cpp:26-28)
    CUC(cudaMemset(c->rho_p, 0, cap * sizeof(float2)));
    CUC(cudaMemset(c->acc, 0, cap * sizeof(float4)));
    CUC(cudaMalloc(&c->sc, sizeof(DeviceScalars)));
    CUC(cudaMemset(c->sc, 0, sizeof(DeviceScalars)));
    CUC(cudaMallocHost(&c->h_sc, sizeof(DeviceScalars)));
    memset(c->h_sc, 0, sizeof(DeviceScalars));
    CUC(cudaMalloc(&c->d_box, 6 * sizeof(int)));
    CUC(cudaMalloc(&c->d_counts, (2 * kMaxRanks + 1) * sizeof(unsigned int)));
#undef CUC
    *out = c;
    return SPHB_OK;
}

void sphb_destroy(sphb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_all(ctx);
    delete ctx;
}

int sphb_set_option(sphb_ctx* c, int option, int64_t value) {
    if (!c) return SPHB_E_INVALID;
    switch (option) {
        case SPHB_OPT_MATH_MODE:
            if (value != 0 && value != 1) return fail(c, SPHB_E_INVALID, "math mode must be 0 (strict) or 1 (fast)");
            c->math_mode = (int)value;
            return SPHB_OK;
        case SPHB_OPT_WALK_RADIUS:
            if (value < 1 || value > 3) return fail(c, SPHB_E_INVALID, "walk radius must be 1..3");
            c->walk_radius = (int)value;
            return SPHB_OK;
        case SPHB_OPT_STAGE_TIMING:
            if (!value) drain_events(c);
            c->stage_timing = value ? 1 : 0;
            return SPHB_OK;
        case SPHB_OPT_DEBUG_CAPTURE:
            c->debug_capture = value ? 1 : 0;
            if (c->debug_capture) { cudaSetDevice(c->device); return ensure_debug(c); }
            return SPHB_OK;
        case SPHB_OPT_PAIR_KERNEL:
            if (value != 0 && value != 2) return fail(c, SPHB_E_INVALID, "pair kernel variant must be 0 (tested walk) or 2 (bitmask hand-off)");
            c->pair_kernel = (int)value;
            return SPHB_OK;
        case SPHB_OPT_GRID_REFINE:
            if (value < 1 || value > 6) return fail(c, SPHB_E_INVALID, "grid refine must be 1..6");
            c->grid_refine = (int)value;
            return SPHB_OK;
        case SPHB_OPT_PAIR_MODE:
            if (value < 0 || value > 2) return fail(c, SPHB_E_INVALID, "pair mode must be 0 (per-lane global loads), 1 (staged) or 2 (staged density, per-lane force)");
            c->pair_mode = (int)value;
            return SPHB_OK;
        case SPHB_OPT_LAYOUT_MAJOR:
            if (value < 0 || value > 2) return fail(c, SPHB_E_INVALID, "layout major axis must be 0..2");
            c->layout_major = (int)value;
            return SPHB_OK;
        case SPHB_OPT_STEP_GRAPHS:
            c->use_graphs = value ? 1 : 0;
            return SPHB_OK;
        case SPHB_OPT_LANES_PER_PARTICLE:
            if (value != 1 && value != 2 && value != 4 && value != 8) return fail(c, SPHB_E_INVALID, "lanes per particle must be 1, 2, 4 or 8");
            c->lanes = (int)value;
            return SPHB_OK;
        case SPHB_OPT_KERNEL_TYPE:
            if (value < 0 || value > 2) return fail(c, SPHB_E_INVALID, "kernel type must be 0 (cubic spline), 1 (Wendland C2) or 2 (Gaussian)");
            c->kernel_type = (int)value;
            return SPHB_OK;
        default:
            return fail(c, SPHB_E_INVALID, "unknown option %d", option);
    }
}

int sphb_get_option(const sphb_ctx* c, int option, int64_t* value) {
    if (!c || !value) return SPHB_E_INVALID;
    switch (option) {
        case SPHB_OPT_MATH_MODE: *value = c->math_mode; return SPHB_OK;
        case SPHB_OPT_WALK_RADIUS: *value = c->walk_radius; return SPHB_OK;
        case SPHB_OPT_STAGE_TIMING: *value = c->stage_timing; return SPHB_OK;
        case SPHB_OPT_DEBUG_CAPTURE: *value = c->debug_capture; return SPHB_OK;
        case SPHB_OPT_PAIR_KERNEL: *value = c->pair_kernel; return SPHB_OK;
        case SPHB_OPT_GRID_REFINE: *value = c->grid_refine; return SPHB_OK;
        case SPHB_OPT_LAYOUT_MAJOR: *value = c->layout_major; return SPHB_OK;
        case SPHB_OPT_PAIR_MODE: *value = c->pair_mode; return SPHB_OK;
        case SPHB_OPT_KERNEL_TYPE: *value = c->kernel_type; return SPHB_OK;
        case SPHB_OPT_STEP_GRAPHS: *value = c->use_graphs; return SPHB_OK;
        case SPHB_OPT_LANES_PER_PARTICLE: *value = c->lanes; return SPHB_OK;
        default: return SPHB_E_INVALID;
    }
}

int sphb_set_stream(sphb_ctx* c, void* cuda_stream) {
    if (!c) return SPHB_E_INVALID;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    return SPHB_OK;
}

int sphb_synchronize(sphb_ctx* c) {
    if (!c) return SPHB_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    return SPHB_OK;
}

int sphb_set_params(sphb_ctx* c, const sphb_params* p) {
    if (!c || !p) return SPHB_E_INVALID;
    c->prm = *p;
    c->have_params = true;
    return SPHB_OK;
}

int sphb_get_params(const sphb_ctx* c, sphb_params* p) {
    if (!c || !p) return SPHB_E_INVALID;
    *p = c->prm;
    return SPHB_OK;
}

int sphb_size(const sphb_ctx* c, size_t* n) {
    if (!c || !n) return SPHB_E_INVALID;
    *n = c->n;
    return SPHB_OK;
}

static int after_upload(sphb_ctx* c, size_t n) {
    c->box_tracking = false;
    c->n = n;
    c->cur = 0;
    c->box_pending = n > 0;
    c->stepped_since_upload = false;
    return SPHB_OK;
}

static int reset_box_and_speed(sphb_ctx* c) {
    // ordered-int identities: min slots start at +max, max slots at -max
    const int init[6] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    CU(c, cudaMemcpyAsync(c->d_box, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemsetAsync(&c->sc->max_v2_bits, 0, sizeof(unsigned), c->stream));
    return SPHB_OK;
}

}  // extern "C"

// bounding box of the uploaded positions, reduced on the device
namespace sphb {
int launch_bbox(size_t n, const float4* posm, int* d_box, cudaStream_t st);
}

extern "C" {

// Staging of an upload: the host-to-device copies do not touch the particle state, so they need not wait for the work that is
// still queued on the context's stream (the previous step, a read-back in flight): they run on a stream of their own
// (c->in_stream) into their own staging buffer (c->d_in), and only the kernel that installs the new state — enqueued by the
// caller on the context's stream after stage_in_commit — is ordered behind both.
static int stage_in_begin(sphb_ctx* c, size_t bytes) {
    if (!c->in_stream) {
        CU(c, cudaStreamCreateWithFlags(&c->in_stream, cudaStreamNonBlocking));
        CU(c, cudaEventCreateWithFlags(&c->ev_in_done, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&c->ev_in_free, cudaEventDisableTiming));
    }
    if (bytes > c->in_bytes) {
        cudaFree(c->d_in);   // (synchronises the device: no copy into the old buffer is in flight afterwards)
        c->d_in = nullptr; c->in_bytes = 0;
        CU(c, cudaMalloc(&c->d_in, bytes));
        c->in_bytes = bytes;
    }
    if (c->in_used) CU(c, cudaStreamWaitEvent(c->in_stream, c->ev_in_free, 0));   // the previous upload's kernels have read the buffer
    return SPHB_OK;
}
static int stage_in_commit(sphb_ctx* c) {    // after the copies were enqueued on c->in_stream
    CU(c, cudaEventRecord(c->ev_in_done, c->in_stream));
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_in_done, 0));
    return SPHB_OK;
}
static int stage_in_release(sphb_ctx* c) {   // after the consuming kernels were enqueued on c->stream
    CU(c, cudaEventRecord(c->ev_in_free, c->stream));
    c->in_used = true;
    return SPHB_OK;
}

int sphb_upload(sphb_ctx* c, size_t n, const float* pos3, const float* vel3, const float* mass) {
    if (!c) return SPHB_E_INVALID;
    if (n > c->capacity) return fail(c, SPHB_E_CAPACITY, "upload of %zu particles exceeds capacity %zu", n, c->capacity);
    if (n > 0 && !pos3) return fail(c, SPHB_E_INVALID, "pos3 is NULL");
    CU(c, cudaSetDevice(c->device));
    if (n == 0) return after_upload(c, 0);
    int rc = stage_in_begin(c, n * 7 * sizeof(float));
    if (rc) return rc;
    float* d_pos = c->d_in;
    float* d_vel = d_pos + 3 * n;
    float* d_mass = d_vel + 3 * n;
    CU(c, cudaMemcpyAsync(d_pos, pos3, n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->in_stream));
    if (vel3) CU(c, cudaMemcpyAsync(d_vel, vel3, n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->in_stream));
    if (mass) CU(c, cudaMemcpyAsync(d_mass, mass, n * sizeof(float), cudaMemcpyHostToDevice, c->in_stream));
    rc = stage_in_commit(c);
    if (rc) return rc;
    rc = reset_box_and_speed(c);
    if (rc) return rc;
    c->stats.kernel_launches += launch_pack_upload(n, d_pos, vel3 ? d_vel : nullptr, mass ? d_mass : nullptr,
                                                   c->prm.particle_mass, c->posm[0], c->velid[0], c->sc, c->stream);
    c->stats.kernel_launches += launch_bbox(n, c->posm[0], c->d_box, c->stream);
    rc = stage_in_release(c);
    if (rc) return rc;
    CU(c, cudaGetLastError());
    return after_upload(c, n);
}

int sphb_upload_strided(sphb_ctx* c, size_t n, const void* base, size_t stride, size_t off_pos, size_t off_vel,
                        size_t off_mass) {
    if (!c) return SPHB_E_INVALID;
    if (n > c->capacity) return fail(c, SPHB_E_CAPACITY, "upload of %zu particles exceeds capacity %zu", n, c->capacity);
    if (n > 0 && !base) return fail(c, SPHB_E_INVALID, "base is NULL");
    if (stride % 4 || off_pos % 4 || off_vel % 4 || off_mass % 4 || off_pos + 12 > stride || off_vel + 12 > stride ||
        off_mass + 4 > stride)
        return fail(c, SPHB_E_INVALID, "stride/offsets must be 4-byte aligned and inside the record");
    CU(c, cudaSetDevice(c->device));
    if (n == 0) return after_upload(c, 0);
    int rc = ensure_stage(c, n * stride);
    if (rc) return rc;
    CU(c, cudaMemcpyAsync(c->d_stage, base, n * stride, cudaMemcpyHostToDevice, c->stream));
    rc = reset_box_and_speed(c);
    if (rc) return rc;
    c->stats.kernel_launches +=
        launch_unpack_strided(n, c->d_stage, stride, off_pos, off_vel, off_mass, c->posm[0], c->velid[0], c->sc, c->stream);
    c->stats.kernel_launches += launch_bbox(n, c->posm[0], c->d_box, c->stream);
    CU(c, cudaGetLastError());
    return after_upload(c, n);
}

int sphb_download(sphb_ctx* c, float* pos3, float* vel3, float* rho, float* pressure, float* acc3) {
    if (!c) return SPHB_E_INVALID;
    if (c->slab_on) return fail(c, SPHB_E_INVALID, "slab mode: ids are global, use sphb_slab_download");
    CU(c, cudaSetDevice(c->device));
    const size_t n = c->n;
    if (n == 0) return SPHB_OK;
    int rc = ensure_stage(c, n * 11 * sizeof(float));
    if (rc) return rc;
    float* d_pos = reinterpret_cast<float*>(c->d_stage);
    float* d_vel = d_pos + 3 * n;
    float* d_acc = d_vel + 3 * n;
    float* d_rho = d_acc + 3 * n;
    float* d_P = d_rho + n;
    // density, pressure and acceleration belong to the particles of the last STEP: after an upload they still sit in the
    // previous set's slot order, so until the next step they read as zeros instead of landing on the wrong particles
    if (!c->stepped_since_upload) {
        if (rho) memset(rho, 0, n * sizeof(float));
        if (pressure) memset(pressure, 0, n * sizeof(float));
        if (acc3) memset(acc3, 0, n * 3 * sizeof(float));
        rho = nullptr; pressure = nullptr; acc3 = nullptr;
        if (!pos3 && !vel3) return SPHB_OK;
    }
    c->stats.kernel_launches += launch_unpermute(n, c->posm[c->cur], c->velid[c->cur], c->rho_p, c->acc, nullptr, nullptr,
                                                 pos3 ? d_pos : nullptr, vel3 ? d_vel : nullptr, rho ? d_rho : nullptr,
                                                 pressure ? d_P : nullptr, acc3 ? d_acc : nullptr, nullptr, nullptr, nullptr,
                                                 c->stream);
    CU(c, cudaGetLastError());
    if (pos3) CU(c, cudaMemcpyAsync(pos3, d_pos, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (vel3) CU(c, cudaMemcpyAsync(vel3, d_vel, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (acc3) CU(c, cudaMemcpyAsync(acc3, d_acc, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (rho) CU(c, cudaMemcpyAsync(rho, d_rho, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (pressure) CU(c, cudaMemcpyAsync(pressure, d_P, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return SPHB_OK;
}

int sphb_download_end(sphb_ctx* c) {
    if (!c) return SPHB_E_INVALID;
    if (!c->out_pending) return SPHB_OK;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaEventSynchronize(c->ev_out_done));
    c->out_pending = false;
    return SPHB_OK;
}

// staging of an asynchronous read-back: waits for the previous one (one transfer in flight: its staging buffer and the
// caller's arrays are reused), then makes sure the stream, the events and `bytes` of staging exist
static int stage_out_begin(sphb_ctx* c, size_t bytes) {
    int rc = sphb_download_end(c);
    if (rc) return rc;
    if (!c->out_stream) {
        CU(c, cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
        CU(c, cudaEventCreateWithFlags(&c->ev_out_ready, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&c->ev_out_done, cudaEventDisableTiming));
    }
    if (bytes > c->out_bytes) {
        cudaFree(c->d_out);
        c->d_out = nullptr; c->out_bytes = 0;
        CU(c, cudaMalloc(&c->d_out, bytes));
        c->out_bytes = bytes;
    }
    return SPHB_OK;
}

int sphb_download_begin(sphb_ctx* c, float* pos3, float* vel3, float* rho, float* pressure, float* acc3) {
    if (!c) return SPHB_E_INVALID;
    if (c->slab_on) return fail(c, SPHB_E_INVALID, "slab mode: ids are global, use sphb_slab_download");
    CU(c, cudaSetDevice(c->device));
    const size_t n = c->n;
    int rc = stage_out_begin(c, n * 11 * sizeof(float));
    if (rc) return rc;
    if (n == 0) return SPHB_OK;
    float* d_pos = c->d_out;
    float* d_vel = d_pos + 3 * n;
    float* d_acc = d_vel + 3 * n;
    float* d_rho = d_acc + 3 * n;
    float* d_P = d_rho + n;
    if (!c->stepped_since_upload) {   // see sphb_download
        if (rho) memset(rho, 0, n * sizeof(float));
        if (pressure) memset(pressure, 0, n * sizeof(float));
        if (acc3) memset(acc3, 0, n * 3 * sizeof(float));
        rho = nullptr; pressure = nullptr; acc3 = nullptr;
        if (!pos3 && !vel3) return SPHB_OK;
    }
    c->stats.kernel_launches += launch_unpermute(n, c->posm[c->cur], c->velid[c->cur], c->rho_p, c->acc, nullptr, nullptr,
                                                 pos3 ? d_pos : nullptr, vel3 ? d_vel : nullptr, rho ? d_rho : nullptr,
                                                 pressure ? d_P : nullptr, acc3 ? d_acc : nullptr, nullptr, nullptr, nullptr,
                                                 c->stream);
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->ev_out_ready, c->stream));
    CU(c, cudaStreamWaitEvent(c->out_stream, c->ev_out_ready, 0));
    if (pos3) CU(c, cudaMemcpyAsync(pos3, d_pos, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
    if (vel3) CU(c, cudaMemcpyAsync(vel3, d_vel, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
    if (acc3) CU(c, cudaMemcpyAsync(acc3, d_acc, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
    if (rho) CU(c, cudaMemcpyAsync(rho, d_rho, n * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
    if (pressure) CU(c, cudaMemcpyAsync(pressure, d_P, n * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
    CU(c, cudaEventRecord(c->ev_out_done, c->out_stream));
    c->out_pending = true;
    return SPHB_OK;
}

int sphb_download_strided(sphb_ctx* c, void* base, size_t stride, size_t off_pos, size_t off_vel, size_t off_density,
                          size_t off_pressure) {
    if (!c) return SPHB_E_INVALID;
    const size_t n = c->n;
    if (n == 0) return SPHB_OK;
    if (!base) return fail(c, SPHB_E_INVALID, "base is NULL");
    const size_t none = (size_t)-1;
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_bounce(c, n * 8 * sizeof(float));
    if (rc) return rc;
    float* h_pos = reinterpret_cast<float*>(c->h_bounce);
    float* h_vel = h_pos + 3 * n;
    float* h_rho = h_vel + 3 * n;
    float* h_P = h_rho + n;
    rc = sphb_download(c, off_pos != none ? h_pos : nullptr, off_vel != none ? h_vel : nullptr,
                       off_density != none ? h_rho : nullptr, off_pressure != none ? h_P : nullptr, nullptr);
    if (rc) return rc;
    unsigned char* b = static_cast<unsigned char*>(base);
    for (size_t i = 0; i < n; ++i) {
        unsigned char* rec = b + i * stride;
        if (off_pos != none) memcpy(rec + off_pos, h_pos + 3 * i, 12);
        if (off_vel != none) memcpy(rec + off_vel, h_vel + 3 * i, 12);
        if (off_density != none) memcpy(rec + off_density, h_rho + i, 4);
        if (off_pressure != none) memcpy(rec + off_pressure, h_P + i, 4);
    }
    return SPHB_OK;
}

int sphb_step(sphb_ctx* c, float dt) {
    if (!c) return SPHB_E_INVALID;
    if (!c->have_params) return fail(c, SPHB_E_INVALID, "sphb_step before sphb_set_params");
    if (c->n == 0) return SPHB_OK;  // SPHEngine::step returns early on an empty system (sph_engine.cpp:94)
    CU(c, cudaSetDevice(c->device));
    int rc = fetch_box(c);
    if (rc) return rc;
    if (c->slab_on && !c->box_tracking) {
        for (int a = 0; a < 3; ++a) { c->box_min[a] = c->slab.box_min[a]; c->box_max[a] = c->slab.box_max[a]; }
    }
    // strict mode keeps the reference's cell size so the device layout IS the reference order; fast mode
    // may sort on a finer grid (fewer candidates per particle) and walk refine x as many cells per axis
    int refine = (c->math_mode == 0) ? 1 : c->grid_refine;
    // Wendland C2 / Gaussian (SPHB_OPT_KERNEL_TYPE): the tested-walk kernels, on the grid that suits them (cells of nsr / 2)
    const int pair_kernel = c->kernel_type != kKernelCubic ? 0 : c->pair_kernel;
    if (c->kernel_type != kKernelCubic && refine > 2) refine = 2;
    // the mask kernels walk at most kMaskMaxRadius cells per axis: with a wider reference walk (SPHB_OPT_WALK_RADIUS 2 =
    // the reference's own 125-cell query) refine only as far as they support
    if (c->math_mode != 0 && pair_kernel == 2 && c->walk_radius * refine > kMaskMaxRadius && c->walk_radius <= kMaskMaxRadius)
        refine = kMaskMaxRadius / c->walk_radius;
    // pair-kernel variant 2 (bitmask hand-off) walks R = walk_radius * refine cells per axis, R in [2, 4], and uses
    // the fast-mode layout.  A refined cell table that would be too large falls back to coarser grids.
    int variant = 0, layout = -1;
    GridDesc g, gc;
    for (;; --refine) {
        variant = (c->math_mode == 0) ? 0 : pair_kernel;
        const int Rw = c->walk_radius * refine;
        if (variant == 2 && (Rw < kMaskMinRadius || Rw > kMaskMaxRadius)) variant = 0;
        layout = (variant == 2) ? (c->slab_on ? c->slab.axis : c->layout_major) : -1;
        rc = make_grid(c, &g, refine, layout, variant == 2 ? Rw : 0);
        if (rc != SPHB_E_GRID || refine == 1) break;
    }
    if (rc) return rc;
    gc = g;
    const bool dbg_ref_sort = c->debug_capture && (refine > 1 || layout >= 0) && !c->slab_on;
    if (dbg_ref_sort) { rc = make_grid(c, &gc, 1); if (rc) return rc; }
    rc = ensure_cell_table(c, g);
    if (rc) return rc;
    if (variant == 2) {
        const size_t cap = c->capacity ? c->capacity : 1;
        const int Rw = c->walk_radius * refine;
        c->mask_stride = (cap + 31) & ~(size_t)31;
        const size_t need = mask_bytes_per_slot(Rw) * c->mask_stride;
        if (need > c->mask_bytes) {
            if (c->masks) { CU(c, cudaStreamSynchronize(c->stream)); cudaFree(c->masks); c->masks = nullptr; c->mask_bytes = 0; }
            CU(c, cudaMalloc(&c->masks, need));
            c->mask_bytes = need;
        }
    }
    // force-pass records: two arrays of 16-byte halves for the 16-bit mask kernels (conflict-free LDS.128 gathers when
    // staged) and the tested-walk kernels (strict mode, variant 0); ONE 32-byte record per particle (one LDG.E.256) for
    // the 64-bit mask kernels of pair_mask_wide.cu
    const PairConsts pk = make_pair_consts(c->prm, c->kernel_type);
    const int mode = (variant == 2 && c->walk_radius * refine >= 4) ? c->pair_mode : 0;
    const bool split = variant == 2 && c->walk_radius * refine >= 4;   // 16-bit mask kernels (pair_mask.cu, pair_stage.cu)
    if (variant == 2 && !split && !c->fab) {
        const size_t cap = c->capacity ? c->capacity : 1;
        CU(c, cudaMalloc(&c->fab, cap * sizeof(ForceRec)));
        CU(c, cudaMemsetAsync(c->fab, 0, cap * sizeof(ForceRec), c->stream));
    }
    if ((variant != 2 || split) && !c->fa) {
        const size_t cap = c->capacity ? c->capacity : 1;
        CU(c, cudaMalloc(&c->fa, cap * sizeof(float4)));
        CU(c, cudaMalloc(&c->fb, cap * sizeof(float4)));
        CU(c, cudaMemsetAsync(c->fa, 0, cap * sizeof(float4), c->stream));
        CU(c, cudaMemsetAsync(c->fb, 0, cap * sizeof(float4), c->stream));
    }
    if (c->debug_capture) { rc = ensure_debug(c); if (rc) return rc; }

    cudaStream_t st = c->stream;
    const size_t n = c->n;
    const IntegrateConsts ic = make_integrate_consts(c->prm);
    uint64_t launches = 0;
    cudaEvent_t* ev = c->stage_timing ? next_event_set(c) : nullptr;
    const bool timing = ev != nullptr;

    const int in = c->cur, outb = c->cur ^ 1;
    int dbg_sorted = -1;
    // after the clamp every position lies inside the AABB (particle.cpp:122-153), so the next step can size its
    // cell table from the bounds without looking at the device.  When that table would be far larger than the
    // particle count (sparse scene, e.g. a drop in a big box) the integrate kernel also reduces the bounding box
    // of the new positions and the next step reads it back (one small synchronising copy) to build a tight table.
    bool track_box = false;
    {
        const float bmin[3] = {c->prm.xmin, c->prm.ymin, c->prm.zmin}, bmax[3] = {c->prm.xmax, c->prm.ymax, c->prm.zmax};
        double cells = 1.0;
        for (int a = 0; a < 3; ++a) {
            double ext = floor((double)(bmax[a] - bmin[a]) * (double)g.inv_cell) + 1.0;
            if (c->slab_on && a == c->slab.axis) {   // only the slab and its ghost layers can be populated on this axis
                const double slab_ext = (double)refine * ((double)c->slab.own_hi - (double)c->slab.own_lo + 2.0 * c->slab.halo_layers);
                if (slab_ext < ext) ext = slab_ext;
            }
            cells *= ext;
        }
        track_box = cells > 8.0 * (double)n;   // slab mode: sphb_slab_append adds the boxes of arriving records
    }
    // Everything the step puts on the stream, in order: clear + counting sort, the pair passes, the fused integration.
    auto enqueue = [&]() -> int {
        if (timing) cudaEventRecord(ev[0], st);
        // counting sort by cell: count (+ the step's dt) -> scan -> scatter -> reorder (ids ascending inside each cell)
        size_t scratch_off = 0;
        const size_t table_bytes = cell_table_bytes(g, &scratch_off);
        CU(c, cudaMemsetAsync(c->cell_start, 0, table_bytes, st));
        launches += launch_cell_count(n, c->posm[in], c->velid[in], g, c->cell_ticket, c->cell_start,
                                      c->debug_capture ? c->refkeys[in] : nullptr, dbg_ref_sort ? c->dbg_keys[0] : nullptr, gc, c->sc,
                                      dt, ic, st);
        launches += launch_scan_exclusive(c->cell_start, (size_t)g.ncells + 1, reinterpret_cast<unsigned char*>(c->cell_start) + scratch_off, st);
        launches += launch_cell_scatter(n, c->cell_ticket, c->cell_start, c->slot_src, st);
        launches += launch_reorder(n, c->slot_src, c->cell_ticket, c->cell_start, c->posm[in], c->velid[in],
                                   c->debug_capture ? c->refkeys[in] : nullptr, c->posm[outb], c->velid[outb],
                                   c->debug_capture ? c->refkeys[outb] : nullptr, st);
        if (timing) cudaEventRecord(ev[1], st);
        if (dbg_ref_sort) {   // debug only: the reference-order permutation, by the same radix sort over (reference cell, id)
            SortBuffers ds;
            ds.keys[0] = c->dbg_keys[0]; ds.keys[1] = c->dbg_keys[1];
            ds.vals[0] = c->dbg_vals[0]; ds.vals[1] = c->dbg_vals[1];
            int o = 0;
            launches += launch_radix_sort_onesweep(ds, n, 0, gc.id_bits + gc.cell_bits, &o, c->sort_scratch, st);
            dbg_sorted = o;
        }

        PairArgs pa{};
        pa.n = n;
        pa.posm = c->posm[outb];
        pa.velid = c->velid[outb];
        pa.cell_start = c->cell_start;
        pa.rho_p = c->rho_p;
        pa.fa = c->fa;
        pa.fb = c->fb;
        pa.acc = c->acc;
        pa.masks = c->masks;
        pa.mask_stride = c->mask_stride;
        pa.fab = c->fab;
        pa.nbr_count = c->debug_capture ? c->nbr_count : nullptr;
        pa.sc = c->sc;
        pa.grid = g;
        pa.k = pk;
        pa.walk_radius = c->walk_radius * refine;
        pa.strict = c->math_mode == 0;
        pa.variant = variant;
        pa.kernel_type = c->kernel_type;
        pa.lanes = split ? c->lanes : 1;
        pa.mode = mode;
        pa.slab_axis = c->slab_on ? c->slab.axis : -1;
        pa.rho_lo = c->slab_on ? c->slab.own_lo - 1 : 0;
        pa.rho_hi = c->slab_on ? c->slab.own_hi + 1 : 0;
        {
            const int ld = launch_density(pa, st);
            if (timing) cudaEventRecord(ev[2], st);
            const int lf = ld < 0 ? ld : launch_force(pa, st);
            if (ld < 0 || lf < 0)
                return fail(c, SPHB_E_CUDA, "the staged pair kernels cannot be configured on this device (%s); set SPHB_OPT_PAIR_MODE 0",
                            cudaGetErrorString(cudaGetLastError()));
            launches += (uint64_t)(ld + lf);
        }
        if (timing) cudaEventRecord(ev[3], st);
        if (track_box) {
            const int init[6] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, (int)0x80000000, (int)0x80000000, (int)0x80000000};
            CU(c, cudaMemcpyAsync(c->d_box, init, sizeof(init), cudaMemcpyHostToDevice, st));
        }
        launches += launch_integrate(n, c->posm[outb], c->velid[outb], c->acc, ic, c->sc, track_box ? c->d_box : nullptr, st);
        if (timing) cudaEventRecord(ev[4], st);
        CU(c, cudaGetLastError());
        return SPHB_OK;
    };
    // Steps whose launches repeat exactly (same particle count, grid, parameters, buffers — every fixed-dt or
    // adaptive-dt step of a resident simulation without per-step read-backs) are replayed from a CUDA graph: one launch
    // call instead of nine, which is what a small scene's step time consists of.  Two graphs, one per parity of the
    // double-buffered arrays; a configuration is captured the second time it comes up in a row, so callers that change
    // something every step (slab mode: the particle count) simply keep launching directly.
    const bool graph_ok = c->use_graphs && !timing && !c->debug_capture && !track_box;
    if (graph_ok) {
        StepKey key;
        key.add(n); key.add(dt); key.add(in); key.add(st); key.add(g); key.add(pk); key.add(ic);
        key.add(variant); key.add(mode); key.add(c->lanes); key.add(refine); key.add(c->walk_radius); key.add(c->math_mode); key.add(c->kernel_type);
        key.add(c->slab_on ? c->slab.axis : -1); key.add(c->slab_on ? c->slab.own_lo : 0); key.add(c->slab_on ? c->slab.own_hi : 0);
        key.add(c->posm[0]); key.add(c->posm[1]); key.add(c->velid[0]); key.add(c->velid[1]); key.add(c->rho_p); key.add(c->acc);
        key.add(c->fa); key.add(c->fb); key.add(c->fab); key.add(c->masks); key.add(c->mask_stride); key.add(c->cell_start);
        key.add(c->cell_ticket); key.add(c->slot_src); key.add(c->sc);
        sphb_ctx::StepGraph& sg = c->graphs[in];
        if (sg.exec && sg.key == key.bytes) {
            CU(c, cudaGraphLaunch(sg.exec, st));
            launches = sg.launches;
        } else if (sg.seen == key.bytes) {
            if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
            // captured on a stream of our own (the caller's may be the legacy default stream, which cannot be captured;
            // capturing executes nothing) and launched into the caller's stream
            if (!c->capture_stream) CU(c, cudaStreamCreateWithFlags(&c->capture_stream, cudaStreamNonBlocking));
            const cudaStream_t user_stream = st;
            st = c->capture_stream;
            CU(c, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            rc = enqueue();
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            st = user_stream;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(c, SPHB_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
            const cudaError_t ie = cudaGraphInstantiate(&sg.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) { sg.exec = nullptr; return fail(c, SPHB_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie)); }
            sg.key = key.bytes;
            sg.launches = launches;
            CU(c, cudaGraphLaunch(sg.exec, st));
        } else {
            sg.seen = key.bytes;
            rc = enqueue();
            if (rc) return rc;
        }
    } else {
        rc = enqueue();
        if (rc) return rc;
    }
    c->cur = outb;
    c->dbg_sorted = dbg_sorted;
    if (dbg_sorted >= 0) c->dbg_id_bits = gc.id_bits;

    c->box_tracking = track_box;
    if (track_box) {
        c->box_pending = true;
    } else if (!c->slab_on) {
        c->box_min[0] = c->prm.xmin; c->box_max[0] = c->prm.xmax;
        c->box_min[1] = c->prm.ymin; c->box_max[1] = c->prm.ymax;
        c->box_min[2] = c->prm.zmin; c->box_max[2] = c->prm.zmax;
    }
    c->stepped_since_upload = true;

    c->step_count++;
    c->stats.steps++;
    c->stats.total_neighbor_queries += n;
    c->stats.kernel_launches += launches;

    return SPHB_OK;
}

int sphb_run_steps(sphb_ctx* c, size_t n, float dt) {
    for (size_t i = 0; i < n; ++i) {
        int rc = sphb_step(c, dt);
        if (rc) return rc;
    }
    return SPHB_OK;
}

int sphb_get_time(sphb_ctx* c, float* current_time, uint64_t* step_count) {
    if (!c) return SPHB_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = read_scalars(c);
    if (rc) return rc;
    if (current_time) *current_time = c->h_sc->time;
    if (step_count) *step_count = c->step_count;
    return SPHB_OK;
}

int sphb_set_time(sphb_ctx* c, float current_time, uint64_t step_count) {
    if (!c) return SPHB_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpyAsync(&c->sc->time, &current_time, sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->step_count = step_count;
    return SPHB_OK;
}

int sphb_cfl_timestep(sphb_ctx* c, float* dt) {
    if (!c || !dt) return SPHB_E_INVALID;
    if (!c->have_params) return fail(c, SPHB_E_INVALID, "sphb_cfl_timestep before sphb_set_params");
    CU(c, cudaSetDevice(c->device));
    c->stats.kernel_launches += launch_cfl_probe(c->sc, make_integrate_consts(c->prm), c->stream);
    int rc = read_scalars(c);
    if (rc) return rc;
    *dt = c->h_sc->dt;
    return SPHB_OK;
}

int sphb_get_stats(sphb_ctx* c, sphb_stats* out) {
    if (!c || !out) return SPHB_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = read_scalars(c);
    if (rc) return rc;
    drain_events(c);
    c->stats.max_neighbors = c->h_sc->max_neighbors;
    c->stats.error_flags = c->h_sc->error_flags;
    *out = c->stats;
    return SPHB_OK;
}

int sphb_reset_stats(sphb_ctx* c) {
    if (!c) return SPHB_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemsetAsync(&c->sc->max_neighbors, 0, 2 * sizeof(unsigned), c->stream));   // max_neighbors, error_flags
    drain_events(c);
    c->stats = sphb_stats{};
    return SPHB_OK;
}

int sphb_set_colors(sphb_ctx* c, size_t n, const float* rgb3) {
    if (!c) return SPHB_E_INVALID;
    if (!rgb3 && n) return fail(c, SPHB_E_INVALID, "rgb3 is NULL");
    if (n > c->capacity) return fail(c, SPHB_E_CAPACITY, "%zu colours exceed the capacity %zu", n, c->capacity);
    CU(c, cudaSetDevice(c->device));
    if (!c->colors && c->capacity) CU(c, cudaMalloc(&c->colors, c->capacity * 3 * sizeof(float)));
    if (n) CU(c, cudaMemcpyAsync(c->colors, rgb3, n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));   // the caller's array may go away
    c->n_colors = n;
    return SPHB_OK;
}

int sphb_export_instances(sphb_ctx* c, float* dst, int dst_on_device, const float* default_rgb) {
    if (!c) return SPHB_E_INVALID;
    if (c->slab_on) return fail(c, SPHB_E_INVALID, "slab mode: ids are global, use sphb_slab_download");
    const size_t n = c->n;
    if (n == 0) return SPHB_OK;
    if (!dst) return fail(c, SPHB_E_INVALID, "dst is NULL");
    CU(c, cudaSetDevice(c->device));
    const float reference_default[3] = {0.0f, 0.5f, 1.0f};   // sph::Particle::color, reference particle.h:35
    const float* rgb = default_rgb ? default_rgb : reference_default;
    float* d_out = dst;
    if (!dst_on_device) {
        int rc = ensure_stage(c, n * 9 * sizeof(float));
        if (rc) return rc;
        d_out = reinterpret_cast<float*>(c->d_stage);
    }
    c->stats.kernel_launches += launch_export_instances(n, c->posm[c->cur], c->velid[c->cur], c->n_colors >= n ? c->colors : nullptr,
                                                        rgb, d_out, c->stream);
    CU(c, cudaGetLastError());
    if (!dst_on_device) {
        CU(c, cudaMemcpyAsync(dst, d_out, n * 9 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return SPHB_OK;
}

int sphb_diagnostics(sphb_ctx* c, double* sum_density, double* kinetic, float* max_speed) {
    if (!c) return SPHB_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemsetAsync(&c->sc->sum_rho, 0, 2 * sizeof(double) + sizeof(unsigned), c->stream));
    c->stats.kernel_launches += launch_diagnostics(c->n, c->posm[c->cur], c->velid[c->cur], c->rho_p, c->sc, c->stream);
    int rc = read_scalars(c);
    if (rc) return rc;
    if (sum_density) *sum_density = c->stepped_since_upload ? c->h_sc->sum_rho : 0.0;   // no densities of this particle set yet
    if (kinetic) *kinetic = c->h_sc->kinetic;
    if (max_speed) {
        float v2;
        memcpy(&v2, &c->h_sc->diag_max_v2_bits, 4);
        *max_speed = sqrtf(v2);
    }
    return SPHB_OK;
}

int sphb_debug_stencil(int radius, int8_t* reach, float* cell_scale) {
    if (!reach || !cell_scale) return SPHB_E_INVALID;
    *cell_scale = kRefinedCellScale;   // make_grid: refined cells are 1 / this times neighbor_search_radius / refine
    const int n = stencil_reach_table(radius, reinterpret_cast<signed char*>(reach));
    return n < 0 ? SPHB_E_INVALID : n;
}

int sphb_debug_dump(sphb_ctx* c, uint64_t* keys, uint32_t* perm, uint32_t* nbr_count) {
    if (!c) return SPHB_E_INVALID;
    if (c->slab_on) return fail(c, SPHB_E_INVALID, "sphb_debug_dump is not available in slab mode (ids are global)");
    if (!c->debug_capture || !c->refkeys[0]) return fail(c, SPHB_E_INVALID, "enable SPHB_OPT_DEBUG_CAPTURE before the step");
    if (!c->stepped_since_upload) return fail(c, SPHB_E_INVALID, "no step has run since the last upload");
    CU(c, cudaSetDevice(c->device));
    const size_t n = c->n;
    if (n == 0) return SPHB_OK;
    int rc = ensure_stage(c, n * 16);
    if (rc) return rc;
    uint64_t* d_keys = reinterpret_cast<uint64_t*>(c->d_stage);
    uint32_t* d_perm = reinterpret_cast<uint32_t*>(d_keys + n);
    uint32_t* d_cnt = d_perm + n;
    c->stats.kernel_launches += launch_unpermute(n, c->posm[c->cur], c->velid[c->cur], c->rho_p, c->acc, c->refkeys[c->cur],
                                                 c->nbr_count, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                 keys ? d_keys : nullptr, perm ? d_perm : nullptr,
                                                 nbr_count ? d_cnt : nullptr, c->stream);
    CU(c, cudaGetLastError());
    if (keys) CU(c, cudaMemcpyAsync(keys, d_keys, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (perm && c->dbg_sorted < 0) CU(c, cudaMemcpyAsync(perm, d_perm, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (nbr_count) CU(c, cudaMemcpyAsync(nbr_count, d_cnt, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (perm && c->dbg_sorted >= 0) {
        // refined layout: the reference-order permutation is the id field of the separately sorted
        // (reference cell, id) keys
        uint64_t* h = static_cast<uint64_t*>(malloc(n * sizeof(uint64_t)));
        if (!h) return fail(c, SPHB_E_NOMEM, "out of host memory");
        cudaError_t e = cudaMemcpy(h, c->dbg_keys[c->dbg_sorted], n * sizeof(uint64_t), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { free(h); return fail(c, SPHB_E_CUDA, "debug perm copy: %s", cudaGetErrorString(e)); }
        const uint64_t mask = (1ull << c->dbg_id_bits) - 1ull;
        for (size_t s = 0; s < n; ++s) perm[s] = (uint32_t)(h[s] & mask);
        free(h);
    }
    return SPHB_OK;
}


// ---- slab decomposition -----------------------------------------------------------------------------

int sphb_set_slab(sphb_ctx* c, const sphb_slab* slab) {
    if (!c) return SPHB_E_INVALID;
    if (!slab) { c->slab_on = false; return SPHB_OK; }
    if (slab->axis < 0 || slab->axis > 2) return fail(c, SPHB_E_INVALID, "slab axis must be 0..2");
    // own_hi == own_lo: a rank that owns no cells (more ranks than the scene has cell layers to share); it never holds particles
    if (slab->own_hi < slab->own_lo) return fail(c, SPHB_E_INVALID, "inverted slab [%d, %d)", slab->own_lo, slab->own_hi);
    if (slab->halo_layers < 2) return fail(c, SPHB_E_INVALID, "halo_layers must be >= 2 (density of the first ghost layer is recomputed locally)");
    if (slab->id_space == 0 || slab->id_space > (1ull << 31)) return fail(c, SPHB_E_INVALID, "id_space must be in [1, 2^31]");
    c->slab = *slab;
    c->slab_on = true;
    return SPHB_OK;
}

int sphb_upload_ids(sphb_ctx* c, size_t n, const float* pos3, const float* vel3, const float* mass, const uint32_t* ids) {
    if (!c) return SPHB_E_INVALID;
    if (n > c->capacity) return fail(c, SPHB_E_CAPACITY, "upload of %zu particles exceeds capacity %zu", n, c->capacity);
    if (n > 0 && (!pos3 || !ids)) return fail(c, SPHB_E_INVALID, "pos3 / ids is NULL");
    CU(c, cudaSetDevice(c->device));
    if (n == 0) return after_upload(c, 0);
    int rc = stage_in_begin(c, n * 8 * sizeof(float));
    if (rc) return rc;
    float* d_pos = c->d_in;
    float* d_vel = d_pos + 3 * n;
    float* d_mass = d_vel + 3 * n;
    uint32_t* d_ids = reinterpret_cast<uint32_t*>(d_mass + n);
    CU(c, cudaMemcpyAsync(d_pos, pos3, n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->in_stream));
    if (vel3) CU(c, cudaMemcpyAsync(d_vel, vel3, n * 3 * sizeof(float), cudaMemcpyHostToDevice, c->in_stream));
    if (mass) CU(c, cudaMemcpyAsync(d_mass, mass, n * sizeof(float), cudaMemcpyHostToDevice, c->in_stream));
    CU(c, cudaMemcpyAsync(d_ids, ids, n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->in_stream));
    rc = stage_in_commit(c);
    if (rc) return rc;
    rc = reset_box_and_speed(c);
    if (rc) return rc;
    c->stats.kernel_launches += launch_pack_upload_ids(n, d_pos, vel3 ? d_vel : nullptr, mass ? d_mass : nullptr, d_ids,
                                                       c->prm.particle_mass, c->posm[0], c->velid[0], c->stream);
    c->stats.kernel_launches += launch_max_speed(n, c->velid[0], c->sc, c->stream);
    c->stats.kernel_launches += launch_bbox(n, c->posm[0], c->d_box, c->stream);
    rc = stage_in_release(c);
    if (rc) return rc;
    CU(c, cudaGetLastError());
    return after_upload(c, n);
}

static int exchange_count_impl(sphb_ctx* c, const int32_t* cuts, int nranks, int my_rank, SlabCuts* sc) {
    if (!c->slab_on) return fail(c, SPHB_E_INVALID, "sphb_set_slab first");
    if (nranks < 1 || nranks > kMaxRanks || my_rank < 0 || my_rank >= nranks) return fail(c, SPHB_E_INVALID, "bad rank layout");
    CU(c, cudaSetDevice(c->device));
    sc->nranks = nranks;
    for (int d = 0; d <= nranks; ++d) sc->cuts[d] = cuts[d];
    const float ref_inv = 1.0f / c->prm.neighbor_search_radius;
    CU(c, cudaMemsetAsync(c->d_counts, 0, (2 * kMaxRanks + 1) * sizeof(unsigned int), c->stream));
    c->stats.kernel_launches += launch_exchange_count(c->n, c->posm[c->cur], c->velid[c->cur], *sc, c->slab.axis, ref_inv,
                                                      c->slab.halo_layers, c->d_counts, c->stream);
    return SPHB_OK;
}

// h[k]: group sizes of this rank (k = 2r: owned by rank r, 2r+1: ghosts for rank r)
static int exchange_split_impl(sphb_ctx* c, const SlabCuts& sc, int my_rank, const unsigned int* h, void* d_out, size_t cap_records) {
    const int nranks = sc.nranks;
    const float ref_inv = 1.0f / c->prm.neighbor_search_radius;
    const int in = c->cur, out = c->cur ^ 1;
    ExchangeOffsets off;
    size_t total = 0, all = 0;
    for (int k = 0; k < 2 * nranks; ++k) {
        off.start[k] = (unsigned int)total;
        if (k != 2 * my_rank) total += h[k];      // key 2*me = particles kept in place
        if ((k & 1) == 0) all += h[k];
    }
    // the context also holds last step's ghosts, which the exchange drops: owned <= n
    if (all > c->n) return fail(c, SPHB_E_INVALID, "exchange counts cover %zu particles, the context holds only %zu", all, c->n);
    if (total > cap_records) return fail(c, SPHB_E_CAPACITY, "%zu exchange records exceed the buffer (%zu records)", total, cap_records);
    if (total > 0 && !d_out) return fail(c, SPHB_E_INVALID, "d_out is NULL");
    CU(c, cudaMemsetAsync(c->d_counts, 0, (2 * kMaxRanks + 1) * sizeof(unsigned int), c->stream));
    c->stats.kernel_launches += launch_exchange_split(c->n, c->posm[in], c->velid[in], sc, c->slab.axis, ref_inv, c->slab.halo_layers,
                                                      my_rank, c->posm[out], c->velid[out], static_cast<float4*>(d_out), off,
                                                      c->d_counts, c->stream);
    CU(c, cudaGetLastError());
    c->cur = out;
    c->n = h[2 * my_rank];
    c->stepped_since_upload = false;
    return SPHB_OK;
}

int sphb_slab_exchange_pack(sphb_ctx* c, const int32_t* cuts, int nranks, int my_rank, void* d_out, size_t cap_records,
                            uint64_t* counts) {
    if (!c || !cuts || !counts) return SPHB_E_INVALID;
    SlabCuts sc;
    int rc = exchange_count_impl(c, cuts, nranks, my_rank, &sc);
    if (rc) return rc;
    unsigned int h[2 * kMaxRanks];
    CU(c, cudaMemcpyAsync(h, c->d_counts, 2 * nranks * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 2 * nranks; ++k) counts[k] = h[k];
    return exchange_split_impl(c, sc, my_rank, h, d_out, cap_records);
}

int sphb_slab_exchange_count(sphb_ctx* c, const int32_t* cuts, int nranks, int my_rank, uint32_t* d_counts) {
    if (!c || !cuts || !d_counts) return SPHB_E_INVALID;
    SlabCuts sc;
    int rc = exchange_count_impl(c, cuts, nranks, my_rank, &sc);
    if (rc) return rc;
    CU(c, cudaMemcpyAsync(d_counts, c->d_counts, 2 * nranks * sizeof(unsigned int), cudaMemcpyDeviceToDevice, c->stream));
    return SPHB_OK;
}

int sphb_slab_exchange_split(sphb_ctx* c, const int32_t* cuts, int nranks, int my_rank, const uint32_t* counts, void* d_out,
                             size_t cap_records) {
    if (!c || !cuts || !counts) return SPHB_E_INVALID;
    if (!c->slab_on) return fail(c, SPHB_E_INVALID, "sphb_set_slab first");
    if (nranks < 1 || nranks > kMaxRanks || my_rank < 0 || my_rank >= nranks) return fail(c, SPHB_E_INVALID, "bad rank layout");
    CU(c, cudaSetDevice(c->device));
    SlabCuts sc;
    sc.nranks = nranks;
    for (int d = 0; d <= nranks; ++d) sc.cuts[d] = cuts[d];
    return exchange_split_impl(c, sc, my_rank, counts, d_out, cap_records);
}

int sphb_slab_append(sphb_ctx* c, const void* d_in, size_t count, int ghost) {
    if (!c) return SPHB_E_INVALID;
    if (count == 0) return SPHB_OK;
    if (!d_in) return fail(c, SPHB_E_INVALID, "d_in is NULL");
    if (c->n + count > c->capacity) return fail(c, SPHB_E_CAPACITY, "append of %zu records overflows capacity %zu (have %zu)", count, c->capacity, c->n);
    CU(c, cudaSetDevice(c->device));
    if (ghost < 0)   // records carry their own ghost flag (one-round exchange)
        c->stats.kernel_launches += launch_slab_append_asis(count, static_cast<const float4*>(d_in), c->posm[c->cur] + c->n,
                                                            c->velid[c->cur] + c->n, c->stream);
    else
        c->stats.kernel_launches += launch_slab_append(count, static_cast<const float4*>(d_in), ghost != 0, c->posm[c->cur] + c->n,
                                                       c->velid[c->cur] + c->n, c->stream);
    if (c->box_tracking) c->stats.kernel_launches += launch_bbox(count, c->posm[c->cur] + c->n, c->d_box, c->stream);
    CU(c, cudaGetLastError());
    c->n += count;
    c->stepped_since_upload = false;
    return SPHB_OK;
}

int sphb_slab_download(sphb_ctx* c, size_t cap, uint32_t* ids, float* pos3, float* vel3, float* rho, float* pressure,
                       float* acc3, size_t* count) {
    if (!c || !ids || !count) return SPHB_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    const size_t n = c->n;
    *count = 0;
    if (n == 0) return SPHB_OK;
    int rc = ensure_stage(c, n * 12 * sizeof(float));
    if (rc) return rc;
    float* d_pos = reinterpret_cast<float*>(c->d_stage);
    float* d_vel = d_pos + 3 * n;
    float* d_acc = d_vel + 3 * n;
    float* d_rho = d_acc + 3 * n;
    float* d_P = d_rho + n;
    uint32_t* d_ids = reinterpret_cast<uint32_t*>(d_P + n);
    unsigned int h = 0;
    CU(c, cudaMemsetAsync(c->d_counts, 0, sizeof(unsigned int), c->stream));
    c->stats.kernel_launches += launch_slab_export(n, c->posm[c->cur], c->velid[c->cur], c->rho_p, c->acc, d_ids, pos3 ? d_pos : nullptr,
                                                   vel3 ? d_vel : nullptr, rho ? d_rho : nullptr, pressure ? d_P : nullptr,
                                                   acc3 ? d_acc : nullptr, c->d_counts, c->stream);
    CU(c, cudaMemcpyAsync(&h, c->d_counts, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (h > cap) return fail(c, SPHB_E_CAPACITY, "%u owned particles exceed the output capacity %zu", h, cap);
    CU(c, cudaMemcpyAsync(ids, d_ids, h * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (pos3) CU(c, cudaMemcpyAsync(pos3, d_pos, (size_t)h * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (vel3) CU(c, cudaMemcpyAsync(vel3, d_vel, (size_t)h * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (acc3) CU(c, cudaMemcpyAsync(acc3, d_acc, (size_t)h * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (rho) CU(c, cudaMemcpyAsync(rho, d_rho, h * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (pressure) CU(c, cudaMemcpyAsync(pressure, d_P, h * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    *count = h;
    return SPHB_OK;
}


// sphb_slab_download in two halves (see sphb_download_begin): the number of owned particles is only known on the device
// when the copies are enqueued, so they move min(cap, particles held incl. halo copies) entries per field; _end waits and
// reports how many of them are owned particles (the compacted front of every array).
int sphb_slab_download_begin(sphb_ctx* c, size_t cap, uint32_t* ids, float* pos3, float* vel3, float* rho, float* pressure, float* acc3) {
    if (!c || !ids) return SPHB_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    const size_t n = c->n;
    int rc = stage_out_begin(c, n * 12 * sizeof(float));
    if (rc) return rc;
    if (!c->h_box) CU(c, cudaMallocHost(&c->h_box, 8 * sizeof(int)));
    c->h_box[6] = 0;
    c->out_cap = cap;
    if (n == 0) return SPHB_OK;
    float* d_pos = c->d_out;
    float* d_vel = d_pos + 3 * n;
    float* d_acc = d_vel + 3 * n;
    float* d_rho = d_acc + 3 * n;
    float* d_P = d_rho + n;
    uint32_t* d_ids = reinterpret_cast<uint32_t*>(d_P + n);
    CU(c, cudaMemsetAsync(c->d_counts, 0, sizeof(unsigned int), c->stream));
    c->stats.kernel_launches += launch_slab_export(n, c->posm[c->cur], c->velid[c->cur], c->rho_p, c->acc, d_ids, pos3 ? d_pos : nullptr,
                                                   vel3 ? d_vel : nullptr, rho ? d_rho : nullptr, pressure ? d_P : nullptr,
                                                   acc3 ? d_acc : nullptr, c->d_counts, c->stream);
    // the owned count goes to pinned host memory by a kernel (word 6 of the read-back block; no copy engine involved)
    c->stats.kernel_launches += sphb::launch_word_to_host(reinterpret_cast<const int*>(c->d_counts), c->h_box + 6, c->stream);
    CU(c, cudaGetLastError());
    CU(c, cudaEventRecord(c->ev_out_ready, c->stream));
    CU(c, cudaStreamWaitEvent(c->out_stream, c->ev_out_ready, 0));
    const size_t m = n < cap ? n : cap;
    CU(c, cudaMemcpyAsync(ids, d_ids, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->out_stream));
    if (pos3) CU(c, cudaMemcpyAsync(pos3, d_pos, m * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
    if (vel3) CU(c, cudaMemcpyAsync(vel3, d_vel, m * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
    if (acc3) CU(c, cudaMemcpyAsync(acc3, d_acc, m * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
    if (rho) CU(c, cudaMemcpyAsync(rho, d_rho, m * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
    if (pressure) CU(c, cudaMemcpyAsync(pressure, d_P, m * sizeof(float), cudaMemcpyDeviceToHost, c->out_stream));
    CU(c, cudaEventRecord(c->ev_out_done, c->out_stream));
    c->out_pending = true;
    return SPHB_OK;
}

int sphb_slab_download_end(sphb_ctx* c, size_t* count) {
    if (!c || !count) return SPHB_E_INVALID;
    int rc = sphb_download_end(c);
    if (rc) return rc;
    const size_t h = c->h_box ? (size_t)(unsigned int)c->h_box[6] : 0;
    *count = h;
    if (h > c->out_cap) return fail(c, SPHB_E_CAPACITY, "%zu owned particles exceed the output capacity %zu", h, c->out_cap);
    return SPHB_OK;
}

// A small device buffer (the gathered table of exchange group sizes) into PINNED host memory by a kernel on the context's
// stream — not by a device-to-host copy, which would queue on the copy engine behind a bulk read-back still in flight
// (sphb_slab_download_begin) and stall the step behind it.  Enqueue only: the caller synchronises the stream.
int sphb_read_small(sphb_ctx* c, const void* d_src, void* h_dst_pinned, size_t bytes) {
    if (!c || !d_src || !h_dst_pinned) return SPHB_E_INVALID;
    if (bytes % 4 || bytes > (1u << 20)) return fail(c, SPHB_E_INVALID, "sphb_read_small: %zu bytes (a multiple of 4, at most 1 MiB)", bytes);
    CU(c, cudaSetDevice(c->device));
    c->stats.kernel_launches += sphb::launch_words_to_host(static_cast<const uint32_t*>(d_src), static_cast<uint32_t*>(h_dst_pinned),
                                                           (unsigned)(bytes / 4), c->stream);
    CU(c, cudaGetLastError());
    return SPHB_OK;
}

int sphb_get_cfl_state(sphb_ctx* c, float* max_v2, float* a0_xyz, int* a0_fresh) {
    if (!c) return SPHB_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    int rc = read_scalars(c);
    if (rc) return rc;
    if (max_v2) memcpy(max_v2, &c->h_sc->max_v2_bits, 4);
    if (a0_xyz) { a0_xyz[0] = c->h_sc->a0[0]; a0_xyz[1] = c->h_sc->a0[1]; a0_xyz[2] = c->h_sc->a0[2]; }
    if (a0_fresh) *a0_fresh = c->h_sc->a0_fresh ? 1 : 0;
    CU(c, cudaMemsetAsync(&c->sc->a0_fresh, 0, sizeof(unsigned), c->stream));
    return SPHB_OK;
}

int sphb_set_cfl_state(sphb_ctx* c, float max_v2, const float* a0_xyz) {
    if (!c || !a0_xyz) return SPHB_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    unsigned bits;
    memcpy(&bits, &max_v2, 4);
    CU(c, cudaMemcpyAsync(&c->sc->max_v2_bits, &bits, 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(&c->sc->a0[0], a0_xyz, 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return SPHB_OK;
}

}  // extern "C"
