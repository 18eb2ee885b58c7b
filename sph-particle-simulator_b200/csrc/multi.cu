// sphb_multi: ONE engine over several GPUs of a node, behind the C ABI (include/sphb.h, "several GPUs in one process").
//
// NEW functionality — the reference is a single-process CPU program.  One host thread drives one
// sphb_ctx per device: the domain is cut into slabs of whole reference cells along one axis (cuts balanced by particle
// count at upload time), and every step runs the one-round exchange of the slab protocol with
// peer-to-peer copies over NVLink instead of NCCL collectives:
//   count   every context counts, per destination, the particles it has to send (owned elsewhere now = migration;
//           inside a neighbour's two halo layers = ghosts) and the counts are copied to pinned host memory — the step's
//           ONE host wait, taken for all devices at once
//   split   every context groups its outgoing 32-byte records by destination
//   copy    cudaMemcpyPeerAsync per (source, destination) pair on the source's stream; events order them against the
//           destination's append of this step and against its reads of the previous step
//   append  + the local step on owned + ghost particles (sphb_step)
// This is the in-process counterpart of sph-particle-simulator_b200/slab.py (one process per GPU, NCCL), built on the same
// per-context entry points (sphb_set_slab, sphb_slab_exchange_count / _split, sphb_slab_append), so an N-device run
// reproduces the single-context run bit for bit in strict mode (tests/test_multi_gpu.py).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sphb_internal.cuh"

// One worker thread per device: the phases of a step (count, split + copies, append + local step) are enqueued on all
// devices at once.  A single host thread needs ~40 us per device to enqueue a local step; with 8 devices and 2 ms steps
// (strong scaling) the last device would start 0.3 ms late every step.
struct DeviceWorkers {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int(int)> job;
    std::vector<int> rc;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;

    void start(int n) {
        rc.assign((size_t)n, 0);
        for (int d = 0; d < n; ++d) threads.emplace_back([this, d] { loop(d); });
    }
    void loop(int d) {
        uint64_t seen = 0;
        for (;;) {
            std::function<int(int)> fn;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
                fn = job;
            }
            const int r = fn(d);
            {
                std::lock_guard<std::mutex> lk(mu);
                rc[(size_t)d] = r;
                if (--pending == 0) cv_done.notify_one();
            }
        }
    }
    // runs fn(d) for every device concurrently and waits; returns the first non-zero result
    int run(const std::function<int(int)>& fn) {
        if (threads.empty()) return 0;
        {
            std::lock_guard<std::mutex> lk(mu);
            job = fn;
            pending = (int)threads.size();
            ++generation;
        }
        cv_go.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
        for (int r : rc) if (r) return r;
        return 0;
    }
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_go.notify_all();
        for (auto& t : threads) t.join();
        threads.clear();
    }
};

struct sphb_multi {
    int ndev = 0;
    std::vector<int> devs;
    std::vector<sphb_ctx*> ctx;
    std::vector<cudaStream_t> streams;
    std::vector<float4*> send, recv;          // exchange buffers (device), xcap records of two float4 each
    std::vector<uint32_t*> d_counts;          // 2 * ndev group sizes per device
    std::vector<cudaEvent_t> ev_sent, ev_appended;
    uint32_t* h_table = nullptr;              // pinned, ndev x 2 ndev
    size_t xcap = 0, cap_ctx = 0, capacity = 0, n_total = 0;
    std::vector<int32_t> cuts;
    int axis = -1;                            // -1: the longest axis of the uploaded particles' bounding box
    int slab_axis = -1;                       // the axis the current cuts are on
    int layers = 2;                           // halo layers per face: 2 = the minimum (SPHB_OPT_MULTI_HALO_LAYERS: wider halos for diagnosis)
    int active = 0;                           // slabs that own cells (the last `active` devices; the others own an empty range)
    sphb_params prm{};
    bool have_params = false, planned = false;
    float a0[3] = {0.0f, 0.0f, 0.0f};
    uint64_t step_count = 0;
    std::vector<uint64_t> owned, ghosts;      // per device, after the last exchange (before the first: the uploaded shares)
    std::vector<float> h_mass;                // per-particle masses as uploaded (empty: the default mass) — needed to re-cut the slabs
    bool want_rebalance = false;              // the last exchange left the devices unbalanced: re-cut before the next step
    uint64_t rebalances = 0;
    uint64_t rebalance_min = 32768;           // ... and at least this many particles above it (SPHB_OPT_MULTI_REBALANCE_MIN)
    std::string err;
    std::vector<std::string> err_dev;         // per device: message of a failure inside a worker
    DeviceWorkers workers;
};

namespace {


int mfail(sphb_multi* m, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (m) m->err = buf;
    return code;
}

int from_ctx(sphb_multi* m, int d, int rc) {
    if (rc != SPHB_OK) {
        const char* e = sphb_last_error(m->ctx[d]);
        m->err = std::string("device ") + std::to_string(m->devs[d]) + ": " + (e ? e : "error");
    }
    return rc;
}

#define MCU(m, call)                                                                                   \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return mfail(m, SPHB_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_));  \
    } while (0)
#define MCTX(m, d, call)                          \
    do {                                          \
        int rc_ = from_ctx(m, d, (call));         \
        if (rc_ != SPHB_OK) return rc_;           \
    } while (0)

// (int)floorf(p * (1.0f / nsr)) in fp32: the reference's cell coordinate (spatial_hash.h:30-36)
inline int ref_cell(float p, float inv) {
    const float f = floorf(p * inv);
    if (!(f == f)) return 0;
    if (f < -2.0e9f) return -2000000000;
    if (f > 2.0e9f) return 2000000000;
    return (int)f;
}

// Cuts of the occupied cell range into nranks slabs of whole cells with near-equal particle counts, every slab at
// least min_width cells wide; the end slabs are open-ended (slab.py plan_cuts).  A scene spanning fewer than
// nranks * min_width cells is shared by as many slabs as fit: the leading devices then own the EMPTY range
// [OPEN_LO, OPEN_LO) — they hold nothing, send nothing and receive nothing — so an engine that owns N GPUs never
// refuses a small scene.  Returns the number of slabs that own cells.
int plan_cuts(const std::vector<int>& cells, int nranks, int min_width, std::vector<int32_t>* cuts) {
    constexpr int32_t kOpenLo = -(1 << 28), kOpenHi = (1 << 28);
    int lo = cells[0], hi = cells[0];
    for (int c : cells) { lo = std::min(lo, c); hi = std::max(hi, c); }
    hi += 1;
    const int active = std::max(1, std::min(nranks, (hi - lo) / min_width));
    std::vector<long long> cum((size_t)(hi - lo) + 1, 0);
    for (int c : cells) cum[(size_t)(c - lo) + 1]++;
    for (size_t k = 1; k < cum.size(); ++k) cum[k] += cum[k - 1];
    const long long total = cum.back();
    cuts->assign((size_t)nranks + 1, kOpenLo);
    (*cuts)[(size_t)nranks] = kOpenHi;
    int prev = lo;
    for (int d = 1; d < active; ++d) {
        const double target = (double)total * d / active;
        int c = lo + (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        if (c > lo && std::fabs((double)cum[(size_t)(c - lo - 1)] - target) < std::fabs((double)cum[(size_t)std::min(c - lo, hi - lo)] - target)) c -= 1;
        c = std::max(c, prev + min_width);
        c = std::min(c, hi - (active - d) * min_width);
        (*cuts)[(size_t)(nranks - active + d)] = c;
        prev = c;
    }
    return active;
}

int rank_of_cell(const std::vector<int32_t>& cuts, int cell) {
    const int g = (int)cuts.size() - 1;
    int r = (int)(std::upper_bound(cuts.begin(), cuts.end(), cell) - cuts.begin()) - 1;
    return std::min(std::max(r, 0), g - 1);
}

// SPHEngine::compute_cfl_timestep (reference sph_engine.cpp:312-333) from globally reduced inputs, in the fp32 operation
// order of the device kernel k_cfl_dt (and of slab.py cfl_timestep): every device then takes the identical dt
float cfl_timestep(const sphb_params& p, float max_v2, const float a0[3]) {
    const float max_velocity = sqrtf(max_v2);
    const float dt_cfl = (p.CFL_factor * p.smoothing_length) / (max_velocity + 1e-6f);
    float a = a0[0] * a0[0] + a0[1] * a0[1];
    a = sqrtf(a + a0[2] * a0[2]);
    const float dt_force = p.CFL_factor * sqrtf(p.smoothing_length / (a + 1e-6f));
    float m = dt_cfl;
    if (dt_force < m) m = dt_force;
    if (p.timestep < m) m = p.timestep;
    return m;
}

void release(sphb_multi* m) {
    for (int d = 0; d < (int)m->ctx.size(); ++d) {
        cudaSetDevice(m->devs[d]);
        if (m->ctx[d]) sphb_destroy(m->ctx[d]);
        if (d < (int)m->send.size()) cudaFree(m->send[d]);
        if (d < (int)m->recv.size()) cudaFree(m->recv[d]);
        if (d < (int)m->d_counts.size()) cudaFree(m->d_counts[d]);
        if (d < (int)m->ev_sent.size() && m->ev_sent[d]) cudaEventDestroy(m->ev_sent[d]);
        if (d < (int)m->ev_appended.size() && m->ev_appended[d]) cudaEventDestroy(m->ev_appended[d]);
        if (d < (int)m->streams.size() && m->streams[d]) cudaStreamDestroy(m->streams[d]);
    }
    if (m->h_table) cudaFreeHost(m->h_table);
}

thread_local std::string g_create_error;

}  // namespace

extern "C" int sphb_multi_upload(sphb_multi* m, size_t n, const float* pos3, const float* vel3, const float* mass);
extern "C" int sphb_multi_download(sphb_multi* m, float* pos3, float* vel3, float* rho, float* pressure, float* acc3);

namespace {

// Re-cut the slabs for the CURRENT positions (one round trip of positions and velocities through the host; rare — only
// when the flow has left one device with 1.5x its even share).  Results do not depend on where the cuts are (every step
// re-sorts owned + halo particles by (cell, global id)); time, step count and the adaptive-dt state are carried over.
int rebalance(sphb_multi* m) {
    const size_t n = m->n_total;
    std::vector<float> p(3 * n), v(3 * n);
    int rc = sphb_multi_download(m, p.data(), v.data(), nullptr, nullptr, nullptr);
    if (rc != SPHB_OK) return rc;
    float t = 0.0f;
    MCTX(m, m->ndev - 1, sphb_get_time(m->ctx[(size_t)m->ndev - 1], &t, nullptr));
    float a0[3] = {m->a0[0], m->a0[1], m->a0[2]};
    for (int d = 0; d < m->ndev; ++d) {   // the acceleration of particle 0 lives on the device that advanced it last
        float v2 = 0.0f, a[3] = {0.0f, 0.0f, 0.0f};
        int fresh = 0;
        MCTX(m, d, sphb_get_cfl_state(m->ctx[d], &v2, a, &fresh));
        if (fresh) { a0[0] = a[0]; a0[1] = a[1]; a0[2] = a[2]; }
    }
    const uint64_t steps = m->step_count, reb = m->rebalances;
    rc = sphb_multi_upload(m, n, p.data(), v.data(), m->h_mass.empty() ? nullptr : m->h_mass.data());
    if (rc != SPHB_OK) return rc;
    for (int d = 0; d < m->ndev; ++d) MCTX(m, d, sphb_set_time(m->ctx[d], t, steps));
    m->step_count = steps;
    m->a0[0] = a0[0]; m->a0[1] = a0[1]; m->a0[2] = a0[2];
    m->rebalances = reb + 1;
    m->want_rebalance = false;
    return SPHB_OK;
}

}  // namespace

extern "C" {

const char* sphb_multi_last_error(const sphb_multi* m) { return m ? m->err.c_str() : g_create_error.c_str(); }

int sphb_create_multi(sphb_multi** out, size_t capacity, int ndev, const int* devices) {
    if (!out) return SPHB_E_INVALID;
    *out = nullptr;
    if (ndev < 1 || ndev > sphb::kMaxRanks || !devices) { g_create_error = "ndev must be 1..64 and devices non-NULL"; return SPHB_E_INVALID; }
    if (capacity >= (1ull << 31)) { g_create_error = "capacity must be below 2^31"; return SPHB_E_CAPACITY; }
    sphb_multi* m = new (std::nothrow) sphb_multi();
    if (!m) return SPHB_E_NOMEM;
    m->ndev = ndev;
    m->capacity = capacity;
    m->devs.assign(devices, devices + ndev);
    // a slab holds its share of the particles, the ghost layers of both faces and what migration adds before the next
    // re-plan: twice the even share plus slack (device memory is ~0.45 kB per unit of capacity)
    m->cap_ctx = ndev == 1 ? std::max<size_t>(capacity, 1) : std::min<size_t>(std::max<size_t>(capacity, 1) + 65536, 2 * (capacity / ndev) + 262144);
    m->xcap = ndev == 1 ? 1 : m->cap_ctx / 2 + 65536;
    m->ctx.assign(ndev, nullptr);
    m->streams.assign(ndev, nullptr);
    m->send.assign(ndev, nullptr);
    m->recv.assign(ndev, nullptr);
    m->d_counts.assign(ndev, nullptr);
    m->ev_sent.assign(ndev, nullptr);
    m->ev_appended.assign(ndev, nullptr);
    m->owned.assign(ndev, 0);
    m->ghosts.assign(ndev, 0);
    auto bail = [&](int code, const std::string& why) {
        g_create_error = why;
        release(m);
        delete m;
        return code;
    };
    for (int d = 0; d < ndev; ++d) {
        int rc = sphb_create(&m->ctx[d], m->cap_ctx, devices[d]);
        if (rc != SPHB_OK) return bail(rc, std::string("sphb_create on device ") + std::to_string(devices[d]) + ": " + sphb_last_error(nullptr));
        if (cudaSetDevice(devices[d]) != cudaSuccess || cudaStreamCreateWithFlags(&m->streams[d], cudaStreamNonBlocking) != cudaSuccess ||
            cudaMalloc(&m->send[d], m->xcap * 2 * sizeof(float4)) != cudaSuccess || cudaMalloc(&m->recv[d], m->xcap * 2 * sizeof(float4)) != cudaSuccess ||
            cudaMalloc(&m->d_counts[d], 2 * ndev * sizeof(uint32_t)) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_sent[d], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->ev_appended[d], cudaEventDisableTiming) != cudaSuccess)
            return bail(SPHB_E_CUDA, std::string("exchange buffers on device ") + std::to_string(devices[d]) + ": " + cudaGetErrorString(cudaGetLastError()));
        sphb_set_stream(m->ctx[d], m->streams[d]);
        for (int e = 0; e < d; ++e) {   // peer access both ways where the hardware offers it (NVLink / NVSwitch); copies work without it too
            if (devices[e] == devices[d]) continue;
            int ok = 0;
            if (cudaDeviceCanAccessPeer(&ok, devices[d], devices[e]) == cudaSuccess && ok) {
                cudaSetDevice(devices[d]); cudaDeviceEnablePeerAccess(devices[e], 0);
                cudaSetDevice(devices[e]); cudaDeviceEnablePeerAccess(devices[d], 0);
                cudaGetLastError();   // "already enabled" is fine
            }
        }
    }
    if (cudaMallocHost(&m->h_table, (size_t)ndev * 2 * ndev * sizeof(uint32_t)) != cudaSuccess) return bail(SPHB_E_CUDA, "pinned count table");
    m->err_dev.assign((size_t)ndev, std::string());
    if (ndev > 1) m->workers.start(ndev);
    *out = m;
    return SPHB_OK;
}

void sphb_destroy_multi(sphb_multi* m) {
    if (!m) return;
    m->workers.shutdown();
    for (int d = 0; d < m->ndev; ++d) { cudaSetDevice(m->devs[d]); if (m->streams[d]) cudaStreamSynchronize(m->streams[d]); }
    release(m);
    delete m;
}

int sphb_multi_device_count(const sphb_multi* m) { return m ? m->ndev : 0; }

int sphb_multi_set_option(sphb_multi* m, int option, int64_t value) {
    if (!m) return SPHB_E_INVALID;
    if (option == SPHB_OPT_MULTI_HALO_LAYERS) {
        if (value < 2 || value > 8) return mfail(m, SPHB_E_INVALID, "halo layers must be 2..8");
        m->layers = (int)value;
        return SPHB_OK;
    }
    if (option == SPHB_OPT_MULTI_REBALANCE_MIN) {
        if (value < 0) return mfail(m, SPHB_E_INVALID, "rebalance threshold must be >= 0");
        m->rebalance_min = (uint64_t)value;
        return SPHB_OK;
    }
    if (option == SPHB_OPT_MULTI_AXIS) {
        if (value < -1 || value > 2) return mfail(m, SPHB_E_INVALID, "slab axis must be -1 (automatic), 0, 1 or 2");
        m->axis = (int)value;
        return SPHB_OK;
    }
    for (int d = 0; d < m->ndev; ++d) MCTX(m, d, sphb_set_option(m->ctx[d], option, value));
    return SPHB_OK;
}

int sphb_multi_set_params(sphb_multi* m, const sphb_params* p) {
    if (!m || !p) return SPHB_E_INVALID;
    for (int d = 0; d < m->ndev; ++d) MCTX(m, d, sphb_set_params(m->ctx[d], p));
    m->prm = *p;
    m->have_params = true;
    return SPHB_OK;
}

int sphb_multi_upload(sphb_multi* m, size_t n, const float* pos3, const float* vel3, const float* mass) {
    if (!m) return SPHB_E_INVALID;
    if (!m->have_params) return mfail(m, SPHB_E_INVALID, "sphb_multi_upload before sphb_multi_set_params (the slabs are cut in units of neighbor_search_radius)");
    if (n > m->capacity) return mfail(m, SPHB_E_CAPACITY, "upload of %zu particles exceeds capacity %zu", n, m->capacity);
    if (n > 0 && !pos3) return mfail(m, SPHB_E_INVALID, "pos3 is NULL");
    m->n_total = n;
    m->planned = false;
    m->a0[0] = m->a0[1] = m->a0[2] = 0.0f;
    if (m->ndev == 1) {
        MCTX(m, 0, sphb_set_slab(m->ctx[0], nullptr));
        MCTX(m, 0, sphb_upload(m->ctx[0], n, pos3, vel3, mass));
        m->owned[0] = n;
        return SPHB_OK;
    }
    std::fill(m->owned.begin(), m->owned.end(), 0);
    std::fill(m->ghosts.begin(), m->ghosts.end(), 0);
    if (n == 0) {
        for (int d = 0; d < m->ndev; ++d) MCTX(m, d, sphb_upload_ids(m->ctx[d], 0, nullptr, nullptr, nullptr, nullptr));
        return SPHB_OK;
    }
    if (mass) { if (mass != m->h_mass.data()) m->h_mass.assign(mass, mass + n); } else m->h_mass.clear();
    m->want_rebalance = false;
    float bmin[3] = {pos3[0], pos3[1], pos3[2]}, bmax[3] = {pos3[0], pos3[1], pos3[2]};
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            const float v = pos3[3 * i + a];
            if (v < bmin[a]) bmin[a] = v;
            if (v > bmax[a]) bmax[a] = v;
        }
    int axis = m->axis;
    if (axis < 0) {
        axis = 0;
        for (int a = 1; a < 3; ++a) if (bmax[a] - bmin[a] > bmax[axis] - bmin[axis]) axis = a;
    }
    m->slab_axis = axis;
    const float inv = 1.0f / m->prm.neighbor_search_radius;
    std::vector<int> cells(n);
    for (size_t i = 0; i < n; ++i) cells[i] = ref_cell(pos3[3 * i + axis], inv);
    m->active = plan_cuts(cells, m->ndev, m->layers, &m->cuts);
    const float lo[3] = {std::min(bmin[0], m->prm.xmin), std::min(bmin[1], m->prm.ymin), std::min(bmin[2], m->prm.zmin)};
    const float hi[3] = {std::max(bmax[0], m->prm.xmax), std::max(bmax[1], m->prm.ymax), std::max(bmax[2], m->prm.zmax)};
    std::vector<std::vector<uint32_t>> mine((size_t)m->ndev);
    for (size_t i = 0; i < n; ++i) mine[(size_t)rank_of_cell(m->cuts, cells[i])].push_back((uint32_t)i);
    std::vector<float> p, v, ms;
    for (int d = 0; d < m->ndev; ++d) {
        const std::vector<uint32_t>& ids = mine[(size_t)d];
        const size_t k = ids.size();
        if (k > m->cap_ctx) return mfail(m, SPHB_E_CAPACITY, "slab %d holds %zu particles, more than a device's share (%zu)", d, k, m->cap_ctx);
        p.resize(3 * k); v.resize(vel3 ? 3 * k : 0); ms.resize(mass ? k : 0);
        for (size_t j = 0; j < k; ++j) {
            const size_t i = ids[j];
            p[3 * j] = pos3[3 * i]; p[3 * j + 1] = pos3[3 * i + 1]; p[3 * j + 2] = pos3[3 * i + 2];
            if (vel3) { v[3 * j] = vel3[3 * i]; v[3 * j + 1] = vel3[3 * i + 1]; v[3 * j + 2] = vel3[3 * i + 2]; }
            if (mass) ms[j] = mass[i];
        }
        sphb_slab sl;
        sl.axis = axis; sl.own_lo = m->cuts[(size_t)d]; sl.own_hi = m->cuts[(size_t)d + 1]; sl.halo_layers = m->layers;
        sl.id_space = std::max<uint64_t>(m->capacity, 2);
        for (int a = 0; a < 3; ++a) { sl.box_min[a] = lo[a]; sl.box_max[a] = hi[a]; }
        MCTX(m, d, sphb_set_slab(m->ctx[d], &sl));
        MCTX(m, d, sphb_upload_ids(m->ctx[d], k, p.data(), vel3 ? v.data() : nullptr, mass ? ms.data() : nullptr, ids.data()));
        m->owned[(size_t)d] = k;
    }
    m->planned = true;
    return SPHB_OK;
}

int sphb_multi_step(sphb_multi* m, float dt) {
    if (!m) return SPHB_E_INVALID;
    if (!m->have_params) return mfail(m, SPHB_E_INVALID, "sphb_multi_step before sphb_multi_set_params");
    if (m->n_total == 0) return SPHB_OK;
    const int G = m->ndev;
    if (G == 1) { MCTX(m, 0, sphb_step(m->ctx[0], dt)); m->step_count++; return SPHB_OK; }
    if (!m->planned) return mfail(m, SPHB_E_INVALID, "no particles uploaded");
    if (m->want_rebalance) {
        const int rc = rebalance(m);
        if (rc != SPHB_OK) return rc;
    }
    const int32_t* cuts = m->cuts.data();
    // what a worker reports: the context's own message, or the CUDA error of a call it made for its device
    auto ctx_fail = [m](int d, int rc) {
        const char* e = sphb_last_error(m->ctx[(size_t)d]);
        m->err_dev[(size_t)d] = std::string("device ") + std::to_string(m->devs[(size_t)d]) + ": " + (e ? e : "error");
        return rc;
    };
    auto cuda_fail = [m](int d, const char* what, cudaError_t e) {
        m->err_dev[(size_t)d] = std::string("device ") + std::to_string(m->devs[(size_t)d]) + ": " + what + ": " + cudaGetErrorString(e);
        return (int)SPHB_E_CUDA;
    };
    auto phase = [m](const std::function<int(int)>& fn) {
        const int rc = m->workers.run(fn);
        if (rc != SPHB_OK)
            for (const std::string& e : m->err_dev) if (!e.empty()) { m->err = e; break; }
        for (std::string& e : m->err_dev) e.clear();
        return rc;
    };
#define WCTX(d, call) do { const int rc_ = (call); if (rc_ != SPHB_OK) return ctx_fail(d, rc_); } while (0)
#define WCU(d, call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(d, #call, e_); } while (0)

    // count: every device counts and copies its row of group sizes out — the step's ONE host wait, taken on all devices at once
    int rc = phase([&](int d) -> int {
        WCU(d, cudaSetDevice(m->devs[(size_t)d]));
        WCTX(d, sphb_slab_exchange_count(m->ctx[(size_t)d], cuts, G, d, m->d_counts[(size_t)d]));
        WCU(d, cudaMemcpyAsync(m->h_table + (size_t)d * 2 * G, m->d_counts[(size_t)d], 2 * G * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                               m->streams[(size_t)d]));
        WCU(d, cudaStreamSynchronize(m->streams[(size_t)d]));
        return SPHB_OK;
    });
    if (rc != SPHB_OK) return rc;
    // layout of the exchange: send[d] holds, for r = 0..G-1, [owned by r (none for r = d)][ghosts for r]; in recv[dst] the block
    // of every source lands behind the previous sources' blocks
    std::vector<size_t> soff((size_t)G * (G + 1), 0), roff((size_t)G * G, 0), rin((size_t)G, 0);
    for (int d = 0; d < G; ++d) {
        const uint32_t* row = m->h_table + (size_t)d * 2 * G;
        size_t off = 0;
        for (int r = 0; r < G; ++r) {
            soff[(size_t)d * (G + 1) + r] = off;
            off += (r == d ? 0u : row[2 * r]) + row[2 * r + 1];
        }
        soff[(size_t)d * (G + 1) + G] = off;
        if (off > m->xcap) return mfail(m, SPHB_E_CAPACITY, "device %d sends %zu records, more than the exchange buffer (%zu)", m->devs[d], off, m->xcap);
    }
    for (int dst = 0; dst < G; ++dst) {
        size_t off = 0;
        for (int src = 0; src < G; ++src) {
            roff[(size_t)src * G + dst] = off;
            off += soff[(size_t)src * (G + 1) + dst + 1] - soff[(size_t)src * (G + 1) + dst];
        }
        rin[(size_t)dst] = off;
        uint64_t own = 0, gh = 0;
        for (int src = 0; src < G; ++src) { own += m->h_table[(size_t)src * 2 * G + 2 * dst]; gh += m->h_table[(size_t)src * 2 * G + 2 * dst + 1]; }
        m->owned[(size_t)dst] = own;
        m->ghosts[(size_t)dst] = gh;
        if (own + gh > m->cap_ctx)
            return mfail(m, SPHB_E_CAPACITY, "device %d would hold %llu particles and halo copies, more than its share (%zu)", m->devs[dst],
                         (unsigned long long)(own + gh), m->cap_ctx);
        if (off > m->xcap) return mfail(m, SPHB_E_CAPACITY, "device %d receives %zu records, more than the exchange buffer (%zu)", m->devs[dst], off, m->xcap);
    }
    {   // load balance: the cuts were planned for the uploaded positions; re-cut when one device holds 1.5x its even share
        uint64_t most = 0;
        for (int d = 0; d < G; ++d) most = std::max(most, m->owned[(size_t)d]);
        const uint64_t even = m->n_total / (uint64_t)std::max(1, m->active);
        if (most * 2 > even * 3 && most > even + m->rebalance_min) m->want_rebalance = true;
    }
    // split + copy: every source groups its records by destination and pushes each block to its destination's receive buffer
    rc = phase([&](int src) -> int {
        WCU(src, cudaSetDevice(m->devs[(size_t)src]));
        WCTX(src, sphb_slab_exchange_split(m->ctx[(size_t)src], cuts, G, src, m->h_table + (size_t)src * 2 * G, m->send[(size_t)src], m->xcap));
        for (int dst = 0; dst < G; ++dst) {
            const size_t cnt = soff[(size_t)src * (G + 1) + dst + 1] - soff[(size_t)src * (G + 1) + dst];
            if (!cnt) continue;
            // recv[dst] may still be read by dst's append of the previous step
            WCU(src, cudaStreamWaitEvent(m->streams[(size_t)src], m->ev_appended[(size_t)dst], 0));
            const float4* from = m->send[(size_t)src] + 2 * soff[(size_t)src * (G + 1) + dst];
            float4* to = m->recv[(size_t)dst] + 2 * roff[(size_t)src * G + dst];
            if (m->devs[(size_t)src] == m->devs[(size_t)dst])
                WCU(src, cudaMemcpyAsync(to, from, cnt * 2 * sizeof(float4), cudaMemcpyDeviceToDevice, m->streams[(size_t)src]));
            else
                WCU(src, cudaMemcpyPeerAsync(to, m->devs[(size_t)dst], from, m->devs[(size_t)src], cnt * 2 * sizeof(float4), m->streams[(size_t)src]));
        }
        WCU(src, cudaEventRecord(m->ev_sent[(size_t)src], m->streams[(size_t)src]));
        return SPHB_OK;
    });
    if (rc != SPHB_OK) return rc;
    // append what arrived (after every source's copies: their records are enqueued, the phase above has returned)
    const bool adaptive = dt <= 0.0f;
    std::vector<float> v2s((size_t)G, 0.0f), a0s((size_t)3 * G, 0.0f);
    std::vector<int> fresh((size_t)G, 0);
    auto append = [&](int dst) -> int {
        WCU(dst, cudaSetDevice(m->devs[(size_t)dst]));
        for (int src = 0; src < G; ++src) WCU(dst, cudaStreamWaitEvent(m->streams[(size_t)dst], m->ev_sent[(size_t)src], 0));
        WCTX(dst, sphb_slab_append(m->ctx[(size_t)dst], m->recv[(size_t)dst], rin[(size_t)dst], -1));
        WCU(dst, cudaEventRecord(m->ev_appended[(size_t)dst], m->streams[(size_t)dst]));
        return SPHB_OK;
    };
    if (adaptive) {   // global max |v|^2, the acceleration of particle 0 from the device that advanced it: one more rendezvous
        rc = phase([&](int d) -> int {
            const int r = append(d);
            if (r != SPHB_OK) return r;
            WCTX(d, sphb_get_cfl_state(m->ctx[(size_t)d], &v2s[(size_t)d], &a0s[(size_t)3 * d], &fresh[(size_t)d]));
            return SPHB_OK;
        });
        if (rc != SPHB_OK) return rc;
        float v2max = 0.0f;
        for (int d = 0; d < G; ++d) {
            if (v2s[(size_t)d] > v2max) v2max = v2s[(size_t)d];
            if (fresh[(size_t)d]) { m->a0[0] = a0s[(size_t)3 * d]; m->a0[1] = a0s[(size_t)3 * d + 1]; m->a0[2] = a0s[(size_t)3 * d + 2]; }
        }
        dt = cfl_timestep(m->prm, v2max, m->a0);
    }
    // ... then the local step
    rc = phase([&](int d) -> int {
        if (!adaptive) {
            const int r = append(d);
            if (r != SPHB_OK) return r;
        }
        WCTX(d, sphb_step(m->ctx[(size_t)d], dt));
        return SPHB_OK;
    });
    if (rc != SPHB_OK) return rc;
#undef WCTX
#undef WCU
    m->step_count++;
    return SPHB_OK;
}

int sphb_multi_synchronize(sphb_multi* m) {
    if (!m) return SPHB_E_INVALID;
    for (int d = 0; d < m->ndev; ++d) MCTX(m, d, sphb_synchronize(m->ctx[d]));
    return SPHB_OK;
}

int sphb_multi_size(sphb_multi* m, size_t* n) {
    if (!m || !n) return SPHB_E_INVALID;
    *n = m->n_total;
    return SPHB_OK;
}

int sphb_multi_download(sphb_multi* m, float* pos3, float* vel3, float* rho, float* pressure, float* acc3) {
    if (!m) return SPHB_E_INVALID;
    if (m->ndev == 1) { MCTX(m, 0, sphb_download(m->ctx[0], pos3, vel3, rho, pressure, acc3)); return SPHB_OK; }
    if (m->n_total == 0) return SPHB_OK;
    std::vector<uint32_t> ids(m->cap_ctx);
    std::vector<float> p(pos3 ? 3 * m->cap_ctx : 0), v(vel3 ? 3 * m->cap_ctx : 0), r(rho ? m->cap_ctx : 0), pr(pressure ? m->cap_ctx : 0),
        ac(acc3 ? 3 * m->cap_ctx : 0);
    size_t seen = 0;
    for (int d = 0; d < m->ndev; ++d) {
        size_t k = 0;
        MCTX(m, d, sphb_slab_download(m->ctx[d], m->cap_ctx, ids.data(), pos3 ? p.data() : nullptr, vel3 ? v.data() : nullptr,
                                      rho ? r.data() : nullptr, pressure ? pr.data() : nullptr, acc3 ? ac.data() : nullptr, &k));
        for (size_t j = 0; j < k; ++j) {
            const size_t i = ids[j];
            if (i >= m->n_total) return mfail(m, SPHB_E_INVALID, "device %d returned particle id %zu of %zu", m->devs[d], i, m->n_total);
            if (pos3) { pos3[3 * i] = p[3 * j]; pos3[3 * i + 1] = p[3 * j + 1]; pos3[3 * i + 2] = p[3 * j + 2]; }
            if (vel3) { vel3[3 * i] = v[3 * j]; vel3[3 * i + 1] = v[3 * j + 1]; vel3[3 * i + 2] = v[3 * j + 2]; }
            if (rho) rho[i] = r[j];
            if (pressure) pressure[i] = pr[j];
            if (acc3) { acc3[3 * i] = ac[3 * j]; acc3[3 * i + 1] = ac[3 * j + 1]; acc3[3 * i + 2] = ac[3 * j + 2]; }
        }
        seen += k;
    }
    if (seen != m->n_total) return mfail(m, SPHB_E_INVALID, "the devices own %zu particles in total, %zu were uploaded", seen, m->n_total);
    return SPHB_OK;
}

int sphb_multi_get_time(sphb_multi* m, float* current_time, uint64_t* step_count) {
    if (!m) return SPHB_E_INVALID;
    // every device that owns cells accumulates the same dt sequence; the last one always owns cells
    MCTX(m, m->ndev - 1, sphb_get_time(m->ctx[(size_t)m->ndev - 1], current_time, nullptr));
    if (step_count) *step_count = m->step_count;
    return SPHB_OK;
}

int sphb_multi_set_time(sphb_multi* m, float current_time, uint64_t step_count) {
    if (!m) return SPHB_E_INVALID;
    for (int d = 0; d < m->ndev; ++d) MCTX(m, d, sphb_set_time(m->ctx[d], current_time, step_count));
    m->step_count = step_count;
    return SPHB_OK;
}

int sphb_multi_get_stats(sphb_multi* m, sphb_stats* out) {
    if (!m || !out) return SPHB_E_INVALID;
    sphb_stats acc{};
    for (int d = 0; d < m->ndev; ++d) {
        sphb_stats s{};
        MCTX(m, d, sphb_get_stats(m->ctx[d], &s));
        acc.total_time = std::max(acc.total_time, s.total_time);
        acc.neighbor_search_time = std::max(acc.neighbor_search_time, s.neighbor_search_time);
        acc.density_computation_time = std::max(acc.density_computation_time, s.density_computation_time);
        acc.force_computation_time = std::max(acc.force_computation_time, s.force_computation_time);
        acc.integration_time = std::max(acc.integration_time, s.integration_time);
        acc.max_neighbors = std::max(acc.max_neighbors, s.max_neighbors);
        acc.kernel_launches += s.kernel_launches;
        acc.error_flags |= s.error_flags;
    }
    acc.steps = m->step_count;
    acc.total_neighbor_queries = m->step_count * m->n_total;
    *out = acc;
    return SPHB_OK;
}

int sphb_multi_reset_stats(sphb_multi* m) {
    if (!m) return SPHB_E_INVALID;
    for (int d = 0; d < m->ndev; ++d) MCTX(m, d, sphb_reset_stats(m->ctx[d]));
    return SPHB_OK;
}

int sphb_multi_diagnostics(sphb_multi* m, double* sum_density, double* kinetic, float* max_speed) {
    if (!m) return SPHB_E_INVALID;
    double sr = 0.0, ke = 0.0;
    float vm = 0.0f;
    for (int d = 0; d < m->ndev; ++d) {
        double a = 0.0, b = 0.0;
        float c = 0.0f;
        MCTX(m, d, sphb_diagnostics(m->ctx[d], &a, &b, &c));   // owned particles only: halo copies belong to another device
        sr += a; ke += b;
        if (c > vm || c != c) vm = c;
    }
    if (sum_density) *sum_density = sr;
    if (kinetic) *kinetic = ke;
    if (max_speed) *max_speed = vm;
    return SPHB_OK;
}

int sphb_multi_run_steps(sphb_multi* m, size_t n, float dt) {
    for (size_t s = 0; s < n; ++s) {
        const int rc = sphb_multi_step(m, dt);
        if (rc != SPHB_OK) return rc;
    }
    return SPHB_OK;
}

int sphb_multi_upload_strided(sphb_multi* m, size_t n, const void* base, size_t stride, size_t off_pos, size_t off_vel, size_t off_mass) {
    if (!m) return SPHB_E_INVALID;
    if (n > 0 && !base) return mfail(m, SPHB_E_INVALID, "base is NULL");
    if (m->ndev == 1) {
        MCTX(m, 0, sphb_set_slab(m->ctx[0], nullptr));
        MCTX(m, 0, sphb_upload_strided(m->ctx[0], n, base, stride, off_pos, off_vel, off_mass));
        m->n_total = n;
        m->owned[0] = n;
        m->a0[0] = m->a0[1] = m->a0[2] = 0.0f;
        return SPHB_OK;
    }
    // the records are dealt to the devices by cell anyway: gather the three fields once on the host
    std::vector<float> p(3 * n), v(3 * n), ms(n);
    const char* b = static_cast<const char*>(base);
    for (size_t i = 0; i < n; ++i) {
        const char* rec = b + i * stride;
        std::memcpy(&p[3 * i], rec + off_pos, 3 * sizeof(float));
        std::memcpy(&v[3 * i], rec + off_vel, 3 * sizeof(float));
        std::memcpy(&ms[i], rec + off_mass, sizeof(float));
    }
    return sphb_multi_upload(m, n, p.data(), v.data(), ms.data());
}

int sphb_multi_download_strided(sphb_multi* m, void* base, size_t stride, size_t off_pos, size_t off_vel, size_t off_density,
                                size_t off_pressure) {
    if (!m) return SPHB_E_INVALID;
    if (m->ndev == 1) { MCTX(m, 0, sphb_download_strided(m->ctx[0], base, stride, off_pos, off_vel, off_density, off_pressure)); return SPHB_OK; }
    const size_t n = m->n_total;
    if (n == 0) return SPHB_OK;
    if (!base) return mfail(m, SPHB_E_INVALID, "base is NULL");
    std::vector<float> p(3 * n), v(3 * n), r(n), pr(n);
    const int rc = sphb_multi_download(m, p.data(), v.data(), r.data(), pr.data(), nullptr);
    if (rc != SPHB_OK) return rc;
    char* b = static_cast<char*>(base);
    for (size_t i = 0; i < n; ++i) {
        char* rec = b + i * stride;
        std::memcpy(rec + off_pos, &p[3 * i], 3 * sizeof(float));
        std::memcpy(rec + off_vel, &v[3 * i], 3 * sizeof(float));
        std::memcpy(rec + off_density, &r[i], sizeof(float));
        std::memcpy(rec + off_pressure, &pr[i], sizeof(float));
    }
    return SPHB_OK;
}

// SPHEngine::compute_cfl_timestep (reference sph_engine.cpp:312-333) as a query: nothing is advanced
int sphb_multi_cfl_timestep(sphb_multi* m, float* dt) {
    if (!m || !dt) return SPHB_E_INVALID;
    if (!m->have_params) return mfail(m, SPHB_E_INVALID, "sphb_multi_cfl_timestep before sphb_multi_set_params");
    if (m->ndev == 1) { MCTX(m, 0, sphb_cfl_timestep(m->ctx[0], dt)); return SPHB_OK; }
    float v2max = 0.0f, a0[3] = {m->a0[0], m->a0[1], m->a0[2]};
    for (int d = 0; d < m->ndev; ++d) {
        float v2 = 0.0f, a[3] = {0.0f, 0.0f, 0.0f};
        int fresh = 0;
        MCTX(m, d, sphb_get_cfl_state(m->ctx[d], &v2, a, &fresh));
        if (v2 > v2max) v2max = v2;
        if (fresh) { a0[0] = a[0]; a0[1] = a[1]; a0[2] = a[2]; m->a0[0] = a[0]; m->a0[1] = a[1]; m->a0[2] = a[2]; }
    }
    *dt = cfl_timestep(m->prm, v2max, a0);
    return SPHB_OK;
}

int sphb_multi_layout(sphb_multi* m, int32_t* cuts, int* axis, uint64_t* owned, uint64_t* ghosts) {
    if (!m) return SPHB_E_INVALID;
    if (cuts) for (int d = 0; d <= m->ndev; ++d) cuts[d] = m->planned ? m->cuts[(size_t)d] : 0;
    if (axis) *axis = m->planned ? m->slab_axis : -1;
    for (int d = 0; d < m->ndev; ++d) {
        if (owned) owned[d] = m->owned[(size_t)d];
        if (ghosts) ghosts[d] = m->ghosts[(size_t)d];
    }
    return SPHB_OK;
}

}  // extern "C"
