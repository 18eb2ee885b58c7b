/* TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
 *
 * Plain-C restatement of the reference engine's per-step hot path (sph::SPHEngine::step and
 * everything below it), written from the reference's behaviour, each function citing the
 * reference file:line it follows.  It is the parity oracle that travels to the GPU box
 * (/root/reference does not exist there); it is pinned bit-exactly against the UNMODIFIED
 * reference compiled IEEE-strict (oracle/_ref/liboracle_strict.so) by tests/test_oracle_pin.py
 * in the build container and, everywhere, against the committed golden vectors under
 * tests/golden/ that were generated from that compiled reference (tests/golden/make_golden.py).
 *
 * Pinning status: PINNED against the compiled reference + golden vectors generated from it.
 * The reference itself ships no tests, fixtures or known-answer vectors, so no
 * upstream golden vectors exist to pin against.
 *
 * Evaluation order matters: all arithmetic is IEEE fp32, evaluated in exactly the association
 * order of the reference's C++ expressions; build with -ffp-contract=off and without
 * -ffast-math (see oracle/Makefile) so the compiler neither fuses nor reassociates.
 *
 * Data layout differs from the reference on purpose (SoA + sorted cell table instead of a 76-byte
 * AoS and an unordered_map) — the *results* are what is restated: the cell table is ordered by the
 * same 63-bit key and looked up by exact key match, which reproduces the map's semantics including
 * the 21-bit wrap of the key (spatial_hash.h:20-27); per-cell lists are in ascending particle
 * index exactly as SpatialHash::build produces them (spatial_hash.cpp:19-24).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    /* SPHParameters, reference sph_engine.h:14-34 (same defaults) */
    float rest_density, gas_constant, viscosity, smoothing_length, particle_mass;
    float timestep, gravity, damping, CFL_factor;
    float xmin, xmax, ymin, ymax, zmin, zmax;
    float neighbor_search_radius;
} params_t;

typedef struct {
    size_t cap, n;
    float *px, *py, *pz, *vx, *vy, *vz, *mass;
    /* engine buffers, capacity-length once initialize() ran (sph_engine.cpp:26-29) */
    float *rho, *P, *ax, *ay, *az;
    int buffers_ready;
    params_t prm;
    /* SpatialHash state (spatial_hash.h:13-16) */
    float cell_size, inv_cell;
    /* Kernel state: CubicSplineKernel (kernels.cpp:12-14, 23-36), WendlandC2Kernel (166-169), GaussianKernel (201-205).
     * ktype: 0 = cubic spline (what SPHEngine constructs), 1 = Wendland C2, 2 = Gaussian — the engine holds a
     * std::unique_ptr<Kernel> and calls W / gradW / laplacianW virtually (sph_engine.cpp:209, 232, 236), so the other
     * two classes behind create_kernel (kernels.cpp:224-236) drop in without any other change. */
    int ktype;
    float kh, kh_sq, sigma, wnorm, gnorm, gssi;
    int initialized;
    float time;
    size_t step_count;
    /* neighbour structures of the last update_neighbor_lists() */
    uint64_t* keys;      /* per particle */
    uint32_t* order;     /* particle indices, stably sorted by key */
    uint64_t* ukeys;     /* unique keys ascending */
    uint32_t* ustart;    /* start offsets into order, n_cells+1 */
    size_t n_cells;
    int n_chunks;
    uint32_t** nbr_buf;  /* per chunk growable list storage */
    size_t* nbr_buf_cap;
    size_t* nbr_off;     /* per particle: offset inside its chunk's buffer */
    uint32_t* nbr_cnt;   /* per particle */
    /* PerformanceStats (sph_engine.h:51-59) */
    double t_total, t_nbr, t_rho, t_force, t_int;
    size_t max_neighbors, total_queries;
} engine_t;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static params_t default_params(void) {
    params_t p;
    p.rest_density = 1000.0f; p.gas_constant = 2000.0f; p.viscosity = 0.001f;
    p.smoothing_length = 0.02f; p.particle_mass = 0.001f; p.timestep = 0.001f;
    p.gravity = -9.81f; p.damping = 0.99f; p.CFL_factor = 0.4f;
    p.xmin = -1.0f; p.xmax = 1.0f; p.ymin = -1.0f; p.ymax = 1.0f; p.zmin = -1.0f; p.zmax = 1.0f;
    p.neighbor_search_radius = 0.04f;
    return p;
}

static void unpack(const float* a, params_t* p) {
    p->rest_density = a[0]; p->gas_constant = a[1]; p->viscosity = a[2]; p->smoothing_length = a[3];
    p->particle_mass = a[4]; p->timestep = a[5]; p->gravity = a[6]; p->damping = a[7]; p->CFL_factor = a[8];
    p->xmin = a[9]; p->xmax = a[10]; p->ymin = a[11]; p->ymax = a[12]; p->zmin = a[13]; p->zmax = a[14];
    p->neighbor_search_radius = a[15];
}
static void pack(const params_t* p, float* a) {
    a[0] = p->rest_density; a[1] = p->gas_constant; a[2] = p->viscosity; a[3] = p->smoothing_length;
    a[4] = p->particle_mass; a[5] = p->timestep; a[6] = p->gravity; a[7] = p->damping; a[8] = p->CFL_factor;
    a[9] = p->xmin; a[10] = p->xmax; a[11] = p->ymin; a[12] = p->ymax; a[13] = p->zmin; a[14] = p->zmax;
    a[15] = p->neighbor_search_radius;
}

/* SpatialHash::set_cell_size, spatial_hash.h:63-66 */
static void set_cell_size(engine_t* e, float c) {
    e->cell_size = c;
    e->inv_cell = 1.0f / c;
}

/* Kernel::Kernel + CubicSplineKernel ctor, kernels.cpp:12-14 and 23-36 (3-D constant only) */
static void make_kernel(engine_t* e, float h) {
    e->kh = h;
    e->kh_sq = h * h;
    e->sigma = 1.0f / ((float)M_PI * h * h * h);
    e->wnorm = 21.0f / (2.0f * (float)M_PI * h * h * h);          /* kernels.cpp:168 */
    e->gssi = 1.0f / (h * h);                                      /* kernels.cpp:203 */
    e->gnorm = 1.0f / powf((float)M_PI * h * h, 1.5f);             /* kernels.cpp:204 (std::pow(float, float)) */
}

/* ---------------------------------------------------------------- smoothing kernel (3-D cubic) */

/* kernels.h:117-119 compute_q;  glm::length = sqrt(x*x + y*y + z*z) left to right */
static inline float len3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

/* CubicSplineKernel::W → W_3d, kernels.cpp:140-143, 58-66 */
static inline float kW(const engine_t* e, float rx, float ry, float rz) {
    if (e->ktype == 2) {   /* GaussianKernel::W, kernels.cpp:207-210 */
        float r_sq = rx * rx + ry * ry + rz * rz;
        return e->gnorm * expf(-r_sq * e->gssi);
    }
    float q = len3(rx, ry, rz) / e->kh;
    if (e->ktype == 1) {   /* WendlandC2Kernel::W, kernels.cpp:171-177 */
        if (q >= 2.0f) return 0.0f;
        float tmp = 1.0f - 0.5f * q;
        return e->wnorm * tmp * tmp * tmp * tmp * (2.0f * q + 1.0f);
    }
    if (q >= 0.0f && q <= 1.0f) {
        return e->sigma * (2.0f / 3.0f - q * q + 0.5f * q * q * q);
    } else if (q > 1.0f && q <= 2.0f) {
        float t = 2.0f - q;
        return e->sigma * (1.0f / 6.0f * t * t * t);
    }
    return 0.0f;
}

/* CubicSplineKernel::gradW → gradW_3d, kernels.cpp:145-148, 96-108 */
static inline void kgradW(const engine_t* e, float rx, float ry, float rz, float* gx, float* gy, float* gz) {
    if (e->ktype == 2) {   /* GaussianKernel::gradW, kernels.cpp:212-216: ((((-2 s) norm) exp) * r */
        float r_sq = rx * rx + ry * ry + rz * rz;
        float exp_term = expf(-r_sq * e->gssi);
        float s = -2.0f * e->gssi * e->gnorm * exp_term;
        *gx = s * rx; *gy = s * ry; *gz = s * rz;
        return;
    }
    if (e->ktype == 1) {   /* WendlandC2Kernel::gradW, kernels.cpp:179-190 */
        float r_len = len3(rx, ry, rz);
        *gx = 0.0f; *gy = 0.0f; *gz = 0.0f;
        if (r_len < 1e-6f) return;
        float q = r_len / e->kh;
        if (q >= 2.0f) return;
        float tmp = 1.0f - 0.5f * q;
        float dW_dq = -5.0f * tmp * tmp * tmp * q;
        float s = e->wnorm * dW_dq, d = r_len * e->kh;
        *gx = s * (rx / d); *gy = s * (ry / d); *gz = s * (rz / d);
        return;
    }
    float q = len3(rx, ry, rz) / e->kh;
    float r_len = len3(rx, ry, rz);
    *gx = 0.0f; *gy = 0.0f; *gz = 0.0f;
    if (r_len < 1e-6f) return;
    float x = 0.0f, y = 0.0f, z = 0.0f;
    if (q >= 0.0f && q <= 1.0f) {
        float s = e->sigma * (-2.0f * q + 1.5f * q * q);
        x = s * (rx / r_len); y = s * (ry / r_len); z = s * (rz / r_len);
    } else if (q > 1.0f && q <= 2.0f) {
        float t = 2.0f - q;
        float s = -e->sigma * (0.5f * t * t);
        x = s * (rx / r_len); y = s * (ry / r_len); z = s * (rz / r_len);
    }
    *gx = x / e->kh; *gy = y / e->kh; *gz = z / e->kh;
}

/* CubicSplineKernel::laplacianW → laplacianW_3d, kernels.cpp:150-153, 130-138 */
static inline float klapW(const engine_t* e, float rx, float ry, float rz) {
    if (e->ktype == 2) {   /* GaussianKernel::laplacianW, kernels.cpp:218-222 */
        float r_sq = rx * rx + ry * ry + rz * rz;
        float exp_term = expf(-r_sq * e->gssi);
        return 2.0f * e->gssi * e->gnorm * exp_term * (2.0f * e->gssi * r_sq - 3.0f);
    }
    float q = len3(rx, ry, rz) / e->kh;
    if (e->ktype == 1) {   /* WendlandC2Kernel::laplacianW, kernels.cpp:192-198 */
        if (q >= 2.0f) return 0.0f;
        float tmp = 1.0f - 0.5f * q;
        return e->wnorm * (5.0f / e->kh_sq) * tmp * tmp * (5.0f * q - 3.0f);
    }
    if (q >= 0.0f && q <= 1.0f) {
        return e->sigma * (-2.0f + 3.0f * q) / e->kh_sq;
    } else if (q > 1.0f && q <= 2.0f) {
        float t = 2.0f - q;
        return e->sigma * t / e->kh_sq;
    }
    return 0.0f;
}

/* ------------------------------------------------------------------------------ spatial hash */

/* SpatialHash::hash_position, spatial_hash.h:20-27 */
static inline uint64_t hash_cell(int x, int y, int z) {
    uint64_t h = 0;
    h |= ((uint64_t)(x & 0x1FFFFF) << 42);
    h |= ((uint64_t)(y & 0x1FFFFF) << 21);
    h |= ((uint64_t)(z & 0x1FFFFF));
    return h;
}

/* SpatialHash::get_grid_coords, spatial_hash.h:30-36 */
static inline void grid_coords(const engine_t* e, float x, float y, float z, int* cx, int* cy, int* cz) {
    *cx = (int)floorf(x * e->inv_cell);
    *cy = (int)floorf(y * e->inv_cell);
    *cz = (int)floorf(z * e->inv_cell);
}

/* stable LSD radix sort of (key, idx) by 63-bit key, 16 bits per pass */
static void sort_by_key(size_t n, const uint64_t* keys, uint32_t* order, uint32_t* tmp) {
    size_t* cnt = (size_t*)malloc(65537 * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    for (int pass = 0; pass < 4; ++pass) {
        int sh = 16 * pass;
        memset(cnt, 0, 65537 * sizeof(size_t));
        for (size_t i = 0; i < n; ++i) cnt[((keys[order[i]] >> sh) & 0xFFFF) + 1]++;
        int trivial = 0;
        for (size_t d = 0; d < 65536; ++d) {
            if (cnt[d + 1] == n) trivial = 1;
            cnt[d + 1] += cnt[d];
        }
        if (trivial) continue;
        for (size_t i = 0; i < n; ++i) tmp[cnt[(keys[order[i]] >> sh) & 0xFFFF]++] = order[i];
        memcpy(order, tmp, n * sizeof(uint32_t));
    }
    free(cnt);
}

/* SpatialHash::build, spatial_hash.cpp:15-25: per-cell lists in ascending particle index */
static void hash_build(engine_t* e) {
    size_t n = e->n;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) {
        int cx, cy, cz;
        grid_coords(e, e->px[i], e->py[i], e->pz[i], &cx, &cy, &cz);
        e->keys[i] = hash_cell(cx, cy, cz);
    }
    uint32_t* tmp = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    sort_by_key(n, e->keys, e->order, tmp);
    free(tmp);
    size_t nc = 0;
    for (size_t s = 0; s < n; ++s) {
        uint64_t k = e->keys[e->order[s]];
        if (s == 0 || k != e->ukeys[nc - 1]) {
            e->ukeys[nc] = k;
            e->ustart[nc] = (uint32_t)s;
            nc++;
        }
    }
    e->ustart[nc] = (uint32_t)n;
    e->n_cells = nc;
}

/* unordered_map::find on the key, spatial_hash.cpp:45 */
static inline long find_cell(const engine_t* e, uint64_t key) {
    size_t lo = 0, hi = e->n_cells;
    while (lo < hi) {
        size_t mid = (lo + hi) >> 1;
        if (e->ukeys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < e->n_cells && e->ukeys[lo] == key) ? (long)lo : -1;
}

static void chunk_push(engine_t* e, int c, size_t* len, uint32_t v) {
    if (*len == e->nbr_buf_cap[c]) {
        size_t nc = e->nbr_buf_cap[c] ? e->nbr_buf_cap[c] * 2 : 4096;
        e->nbr_buf[c] = (uint32_t*)realloc(e->nbr_buf[c], nc * sizeof(uint32_t));
        e->nbr_buf_cap[c] = nc;
    }
    e->nbr_buf[c][(*len)++] = v;
}

static inline int chunk_of(const engine_t* e, size_t i) {
    size_t per = (e->n + (size_t)e->n_chunks - 1) / (size_t)e->n_chunks;
    return (int)(i / (per ? per : 1));
}

/* SPHEngine::update_neighbor_lists (sph_engine.cpp:335-353) = build + query_squared per particle
 * (spatial_hash.cpp:31-57): cell_radius = ceil(sqrt(r2) * inv_cell) + 1; cells visited dx, dy, dz
 * ascending; inside a cell ascending particle index; keep j when dot(d, d) <= r2, self included. */
static void update_neighbor_lists(engine_t* e) {
    hash_build(e);
    size_t n = e->n;
    float r2 = e->prm.neighbor_search_radius * e->prm.neighbor_search_radius;
    int R = (int)ceilf(sqrtf(r2) * e->inv_cell) + 1;
    size_t per = (n + (size_t)e->n_chunks - 1) / (size_t)e->n_chunks;
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < e->n_chunks; ++c) {
        size_t lo = (size_t)c * per, hi = lo + per;
        if (hi > n) hi = n;
        size_t len = 0;
        for (size_t i = lo; i < hi; ++i) {
            float xi = e->px[i], yi = e->py[i], zi = e->pz[i];
            int cx, cy, cz;
            grid_coords(e, xi, yi, zi, &cx, &cy, &cz);
            size_t begin = len;
            for (int dx = -R; dx <= R; ++dx)
                for (int dy = -R; dy <= R; ++dy)
                    for (int dz = -R; dz <= R; ++dz) {
                        long cell = find_cell(e, hash_cell(cx + dx, cy + dy, cz + dz));
                        if (cell < 0) continue;
                        for (uint32_t s = e->ustart[cell]; s < e->ustart[cell + 1]; ++s) {
                            uint32_t j = e->order[s];
                            /* within_radius_squared, spatial_hash.h:70-73 */
                            float ddx = xi - e->px[j], ddy = yi - e->py[j], ddz = zi - e->pz[j];
                            if (ddx * ddx + ddy * ddy + ddz * ddz <= r2) chunk_push(e, c, &len, j);
                        }
                    }
            e->nbr_off[i] = begin;
            e->nbr_cnt[i] = (uint32_t)(len - begin);
        }
    }
    /* perf counters (sph_engine.cpp:350-351) — the reference updates them racily; the race-free
     * values are N per step and the true running maximum. */
    e->total_queries += n;
    for (size_t i = 0; i < n; ++i)
        if (e->nbr_cnt[i] > e->max_neighbors) e->max_neighbors = e->nbr_cnt[i];
}

static inline const uint32_t* nbr_list(const engine_t* e, size_t i) {
    return e->nbr_buf[chunk_of(e, i)] + e->nbr_off[i];
}

/* ----------------------------------------------------------------------------- SPH equations */

/* compute_densities + equations::compute_density, sph_engine.cpp:203-212, 370-383 */
static void compute_densities(engine_t* e) {
    size_t n = e->n;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        const uint32_t* l = nbr_list(e, i);
        float density = e->mass[i] * kW(e, 0.0f, 0.0f, 0.0f);
        for (uint32_t k = 0; k < e->nbr_cnt[i]; ++k) {
            uint32_t j = l[k];
            if (j == i) continue;
            density += e->mass[j] * kW(e, e->px[i] - e->px[j], e->py[i] - e->py[j], e->pz[i] - e->pz[j]);
        }
        e->rho[i] = density;
    }
}

/* compute_pressures + equations::compute_pressure, sph_engine.cpp:214-224, 385-388 */
static void compute_pressures(engine_t* e) {
    size_t n = e->n;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) e->P[i] = e->prm.gas_constant * (e->rho[i] - e->prm.rest_density);
}

/* compute_forces (sph_engine.cpp:226-244) with compute_pressure_force (390-412),
 * compute_viscosity_force (414-432), compute_external_forces (434-436), compute_acceleration (438-443) */
static void compute_forces(engine_t* e) {
    size_t n = e->n;
    const float mu = e->prm.viscosity;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        const uint32_t* l = nbr_list(e, i);
        float fpx = 0.0f, fpy = 0.0f, fpz = 0.0f;
        for (uint32_t k = 0; k < e->nbr_cnt[i]; ++k) {
            uint32_t j = l[k];
            if (j == i) continue;
            float rx = e->px[i] - e->px[j], ry = e->py[i] - e->py[j], rz = e->pz[i] - e->pz[j];
            float r_len = len3(rx, ry, rz);
            if (r_len < 1e-6f) continue;
            float pressure_term = (e->P[i] + e->P[j]) / (2.0f * e->rho[j]);
            float gx, gy, gz;
            kgradW(e, rx, ry, rz, &gx, &gy, &gz);
            float s = e->mass[j] * pressure_term;
            fpx -= s * gx; fpy -= s * gy; fpz -= s * gz;
        }
        float fvx = 0.0f, fvy = 0.0f, fvz = 0.0f;
        for (uint32_t k = 0; k < e->nbr_cnt[i]; ++k) {
            uint32_t j = l[k];
            if (j == i) continue;
            float rx = e->px[i] - e->px[j], ry = e->py[i] - e->py[j], rz = e->pz[i] - e->pz[j];
            float ux = e->vx[j] - e->vx[i], uy = e->vy[j] - e->vy[i], uz = e->vz[j] - e->vz[i];
            float lap = klapW(e, rx, ry, rz);
            float s = (e->mass[j] / e->rho[j]) * mu;
            fvx += (s * ux) * lap; fvy += (s * uy) * lap; fvz += (s * uz) * lap;
        }
        float m = e->mass[i];
        float ex = 0.0f * m, ey = e->prm.gravity * m, ez = 0.0f * m;
        e->ax[i] = (fpx + fvx + ex) / m;
        e->ay[i] = (fpy + fvy + ey) / m;
        e->az[i] = (fpz + fvz + ez) / m;
    }
}

/* integrate_leapfrog, sph_engine.cpp:290-310 */
static void integrate_leapfrog(engine_t* e, float dt) {
    size_t n = e->n;
    const float damping = e->prm.damping;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        float hx = 0.5f * e->ax[i] * dt, hy = 0.5f * e->ay[i] * dt, hz = 0.5f * e->az[i] * dt;
        e->vx[i] += hx; e->vy[i] += hy; e->vz[i] += hz;
        e->px[i] += e->vx[i] * dt; e->py[i] += e->vy[i] * dt; e->pz[i] += e->vz[i] * dt;
        e->vx[i] += hx; e->vy[i] += hy; e->vz[i] += hz;
        e->vx[i] *= damping; e->vy[i] *= damping; e->vz[i] *= damping;
    }
}

/* ParticleSystem::apply_boundary_conditions, particle.cpp:122-153 */
static inline void clamp_axis(float* p, float* v, float lo, float hi) {
    const float damping = 0.8f;
    if (*p < lo) { *p = lo; *v *= -damping; }
    else if (*p > hi) { *p = hi; *v *= -damping; }
}
static void apply_boundary_conditions(engine_t* e) {
    size_t n = e->n;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        clamp_axis(&e->px[i], &e->vx[i], e->prm.xmin, e->prm.xmax);
        clamp_axis(&e->py[i], &e->vy[i], e->prm.ymin, e->prm.ymax);
        clamp_axis(&e->pz[i], &e->vz[i], e->prm.zmin, e->prm.zmax);
    }
}

/* compute_cfl_timestep, sph_engine.cpp:312-333 (force criterion reads particle 0 only) */
static float compute_cfl_timestep(const engine_t* e) {
    float max_velocity = 0.0f;
    for (size_t i = 0; i < e->n; ++i) {
        float v = len3(e->vx[i], e->vy[i], e->vz[i]);
        if (max_velocity < v) max_velocity = v;   /* std::max(a, b) = (a < b) ? b : a */
    }
    float dt_cfl = e->prm.CFL_factor * e->prm.smoothing_length / (max_velocity + 1e-6f);
    float dt_force = e->prm.CFL_factor * sqrtf(e->prm.smoothing_length / (len3(e->ax[0], e->ay[0], e->az[0]) + 1e-6f));
    float m = dt_cfl;
    if (dt_force < m) m = dt_force;
    if (e->prm.timestep < m) m = e->prm.timestep;
    return m;
}

/* ------------------------------------------------------------------ generators (host side) */

typedef struct { float* pos3; float* mass; size_t cap, n; } sink_t;
static void sink_put(sink_t* s, float x, float y, float z, float m) {
    if (s->n < s->cap) {
        if (s->pos3) { s->pos3[3 * s->n] = x; s->pos3[3 * s->n + 1] = y; s->pos3[3 * s->n + 2] = z; }
        if (s->mass) s->mass[s->n] = m;
    }
    s->n++;
}

/* create_fluid_block, particle.cpp:166-188 */
static void gen_fluid_block(sink_t* s, const float* c, const float* sz, float spacing, float m) {
    int nx = (int)(sz[0] / spacing), ny = (int)(sz[1] / spacing), nz = (int)(sz[2] / spacing);
    float sx = c[0] - sz[0] * 0.5f, sy = c[1] - sz[1] * 0.5f, szz = c[2] - sz[2] * 0.5f;
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k)
                sink_put(s, sx + (float)i * spacing, sy + (float)j * spacing, szz + (float)k * spacing, m);
}

/* create_boundary_box, particle.cpp:190-229 */
static void gen_face(sink_t* s, const float* start, const float* sz, float spacing, float m, int a1, int a2, int a3, float value) {
    int n1 = (int)(sz[a1] / spacing), n2 = (int)(sz[a2] / spacing);
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j) {
            float p[3] = {start[0], start[1], start[2]};
            p[a1] = start[a1] + (float)i * spacing;
            p[a2] = start[a2] + (float)j * spacing;
            p[a3] = value;
            sink_put(s, p[0], p[1], p[2], m);
        }
}
static void gen_boundary_box(sink_t* s, const float* c, const float* sz, float spacing, float m) {
    float start[3] = {c[0] - sz[0] * 0.5f, c[1] - sz[1] * 0.5f, c[2] - sz[2] * 0.5f};
    gen_face(s, start, sz, spacing, m, 0, 1, 2, start[2]);
    gen_face(s, start, sz, spacing, m, 0, 1, 2, start[2] + sz[2]);
    gen_face(s, start, sz, spacing, m, 1, 2, 0, start[0]);
    gen_face(s, start, sz, spacing, m, 1, 2, 0, start[0] + sz[0]);
    gen_face(s, start, sz, spacing, m, 0, 2, 1, start[1]);
    gen_face(s, start, sz, spacing, m, 0, 2, 1, start[1] + sz[1]);
}

/* utils::create_dam_break_setup, sph_engine.cpp:450-487 (walls first, then fluid) */
static void gen_dam_break(sink_t* s, const float* dam, const float* fluid, float spacing, float m) {
    float bc[3] = {0.0f, dam[1] / 2.0f, 0.0f};
    gen_boundary_box(s, bc, dam, spacing, m);
    float fc[3] = {-dam[0] / 2.0f + fluid[0] / 2.0f, fluid[1] / 2.0f, 0.0f};
    gen_fluid_block(s, fc, fluid, spacing, m);
}

/* utils::create_fluid_drop_setup, sph_engine.cpp:489-514 */
static void gen_fluid_drop(sink_t* s, const float* c, float radius, float spacing, float m) {
    int npd = (int)(2.0f * radius / spacing);
    float sx = c[0] - radius, sy = c[1] - radius, sz = c[2] - radius;
    for (int i = 0; i < npd; ++i)
        for (int j = 0; j < npd; ++j)
            for (int k = 0; k < npd; ++k) {
                float x = sx + (float)i * spacing, y = sy + (float)j * spacing, z = sz + (float)k * spacing;
                float tx = x - c[0], ty = y - c[1], tz = z - c[2];
                if (tx * tx + ty * ty + tz * tz <= radius * radius) sink_put(s, x, y, z, m);
            }
}

/* utils::create_granular_flow_setup, sph_engine.cpp:516-552 */
static void gen_granular(sink_t* s, const float* pile, const float* domain, float spacing, float m) {
    float bc[3] = {0.0f, domain[1] / 2.0f, 0.0f};
    gen_boundary_box(s, bc, domain, spacing, m);
    float gc[3] = {0.0f, pile[1] / 2.0f + spacing, 0.0f};
    gen_fluid_block(s, gc, pile, spacing, m);
}

/* ------------------------------------------------------------------------------- engine API */

/* ParticleSystem::add_particle, particle.cpp:26-33: silently capped at capacity */
static void add_one(engine_t* e, float x, float y, float z, float vx, float vy, float vz, float m) {
    if (e->n >= e->cap) return;
    size_t i = e->n++;
    e->px[i] = x; e->py[i] = y; e->pz[i] = z;
    e->vx[i] = vx; e->vy[i] = vy; e->vz[i] = vz;
    e->mass[i] = m;
}

static void add_generated(engine_t* e, void (*fill)(sink_t*, const void*), const void* arg) {
    sink_t count = {NULL, NULL, 0, 0};
    fill(&count, arg);
    size_t n = count.n;
    float* pos = (float*)malloc((n ? n : 1) * 3 * sizeof(float));
    float* mass = (float*)malloc((n ? n : 1) * sizeof(float));
    sink_t s = {pos, mass, n, 0};
    fill(&s, arg);
    for (size_t i = 0; i < n; ++i) add_one(e, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 0.0f, 0.0f, 0.0f, mass[i]);
    free(pos);
    free(mass);
}

void* port_create(size_t max_particles) {
    engine_t* e = (engine_t*)calloc(1, sizeof(engine_t));
    size_t c = max_particles ? max_particles : 1;
    e->cap = max_particles;
    e->px = (float*)calloc(c, 4); e->py = (float*)calloc(c, 4); e->pz = (float*)calloc(c, 4);
    e->vx = (float*)calloc(c, 4); e->vy = (float*)calloc(c, 4); e->vz = (float*)calloc(c, 4);
    e->mass = (float*)calloc(c, 4);
    e->rho = (float*)calloc(c, 4); e->P = (float*)calloc(c, 4);
    e->ax = (float*)calloc(c, 4); e->ay = (float*)calloc(c, 4); e->az = (float*)calloc(c, 4);
    e->keys = (uint64_t*)calloc(c, 8); e->order = (uint32_t*)calloc(c, 4);
    e->ukeys = (uint64_t*)calloc(c, 8); e->ustart = (uint32_t*)calloc(c + 1, 4);
    e->nbr_off = (size_t*)calloc(c, sizeof(size_t)); e->nbr_cnt = (uint32_t*)calloc(c, 4);
    e->n_chunks = 256;
    e->nbr_buf = (uint32_t**)calloc((size_t)e->n_chunks, sizeof(uint32_t*));
    e->nbr_buf_cap = (size_t*)calloc((size_t)e->n_chunks, sizeof(size_t));
    /* SPHEngine ctor, sph_engine.cpp:13-18: hash cell = default h until initialize() */
    e->prm = default_params();
    set_cell_size(e, e->prm.smoothing_length);
    make_kernel(e, e->prm.smoothing_length);
    return e;
}

void port_destroy(void* h) {
    engine_t* e = (engine_t*)h;
    if (!e) return;
    free(e->px); free(e->py); free(e->pz); free(e->vx); free(e->vy); free(e->vz); free(e->mass);
    free(e->rho); free(e->P); free(e->ax); free(e->ay); free(e->az);
    free(e->keys); free(e->order); free(e->ukeys); free(e->ustart); free(e->nbr_off); free(e->nbr_cnt);
    for (int c = 0; c < e->n_chunks; ++c) free(e->nbr_buf[c]);
    free(e->nbr_buf); free(e->nbr_buf_cap);
    free(e);
}

int port_sizeof_particle(void) { return 76; /* reference particle.h:17-49 */ }
int port_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void port_default_params(float* p16) { params_t p = default_params(); pack(&p, p16); }

void port_reset_stats(void* h) {
    engine_t* e = (engine_t*)h;
    e->t_total = e->t_nbr = e->t_rho = e->t_force = e->t_int = 0.0;
    e->max_neighbors = 0; e->total_queries = 0;
}

/* SPHEngine::initialize, sph_engine.cpp:20-33 */
void port_initialize(void* h, const float* p16) {
    engine_t* e = (engine_t*)h;
    unpack(p16, &e->prm);
    set_cell_size(e, e->prm.neighbor_search_radius);
    make_kernel(e, e->prm.smoothing_length);
    e->ktype = 0;
    e->buffers_ready = 1;
    e->initialized = 1;
    port_reset_stats(h);
}

/* SPHEngine::set_smoothing_length, sph_engine.cpp:158-163 */
void port_set_smoothing_length(void* h, float hh) {
    engine_t* e = (engine_t*)h;
    e->prm.smoothing_length = hh;
    e->prm.neighbor_search_radius = 2.0f * hh;
    set_cell_size(e, e->prm.neighbor_search_radius);
    make_kernel(e, hh);
    e->ktype = 0;
}

/* SPHEngine::set_parameters, sph_engine.cpp:152-156 */
void port_set_parameters(void* h, const float* p16) {
    engine_t* e = (engine_t*)h;
    unpack(p16, &e->prm);
    set_cell_size(e, e->prm.neighbor_search_radius);
    port_set_smoothing_length(h, e->prm.smoothing_length);
}
void port_get_parameters(void* h, float* p16) { pack(&((engine_t*)h)->prm, p16); }
void port_set_gravity(void* h, float g) { ((engine_t*)h)->prm.gravity = g; }
void port_set_viscosity(void* h, float mu) { ((engine_t*)h)->prm.viscosity = mu; }
void port_set_boundaries(void* h, float x0, float x1, float y0, float y1, float z0, float z1) {
    engine_t* e = (engine_t*)h;
    e->prm.xmin = x0; e->prm.xmax = x1; e->prm.ymin = y0; e->prm.ymax = y1; e->prm.zmin = z0; e->prm.zmax = z1;
}

typedef struct { float a[3], b[3], spacing, m, radius; } gen_arg_t;
static void fill_dam(sink_t* s, const void* v) { const gen_arg_t* g = (const gen_arg_t*)v; gen_dam_break(s, g->a, g->b, g->spacing, g->m); }
static void fill_drop(sink_t* s, const void* v) { const gen_arg_t* g = (const gen_arg_t*)v; gen_fluid_drop(s, g->a, g->radius, g->spacing, g->m); }
static void fill_gran(sink_t* s, const void* v) { const gen_arg_t* g = (const gen_arg_t*)v; gen_granular(s, g->a, g->b, g->spacing, g->m); }
static void fill_block(sink_t* s, const void* v) { const gen_arg_t* g = (const gen_arg_t*)v; gen_fluid_block(s, g->a, g->b, g->spacing, g->m); }
static void fill_box(sink_t* s, const void* v) { const gen_arg_t* g = (const gen_arg_t*)v; gen_boundary_box(s, g->a, g->b, g->spacing, g->m); }

static void auto_init(engine_t* e) {
    if (!e->initialized) {
        float p16[16];
        port_default_params(p16);
        port_initialize(e, p16);
    }
}

/* initialize_dam_break / fluid_drop / granular_flow, sph_engine.cpp:35-81
 * (particles_.clear() does not reset time / step count) */
void port_initialize_dam_break(void* h) {
    engine_t* e = (engine_t*)h;
    auto_init(e);
    gen_arg_t g = {{0.4f, 0.6f, 0.8f}, {0.2f, 0.4f, 0.8f}, 0.01f, e->prm.particle_mass, 0.0f};
    e->n = 0;
    add_generated(e, fill_dam, &g);
}
void port_initialize_fluid_drop(void* h) {
    engine_t* e = (engine_t*)h;
    auto_init(e);
    gen_arg_t g = {{0.0f, 0.5f, 0.0f}, {0, 0, 0}, 0.008f, e->prm.particle_mass, 0.1f};
    e->n = 0;
    add_generated(e, fill_drop, &g);
}
void port_initialize_granular_flow(void* h) {
    engine_t* e = (engine_t*)h;
    auto_init(e);
    gen_arg_t g = {{0.3f, 0.4f, 0.8f}, {1.0f, 1.0f, 1.0f}, 0.012f, e->prm.particle_mass, 0.0f};
    e->n = 0;
    add_generated(e, fill_gran, &g);
}

/* SPHEngine::clear_particles, sph_engine.cpp:87-91 */
void port_clear_particles(void* h) {
    engine_t* e = (engine_t*)h;
    e->n = 0; e->time = 0.0f; e->step_count = 0;
}

/* default Particle: velocity 0, mass 1.0 (particle.h:38-42) */
void port_add_particles(void* h, size_t n, const float* pos3, const float* vel3, const float* mass) {
    engine_t* e = (engine_t*)h;
    for (size_t i = 0; i < n; ++i)
        add_one(e, pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2],
                vel3 ? vel3[3 * i] : 0.0f, vel3 ? vel3[3 * i + 1] : 0.0f, vel3 ? vel3[3 * i + 2] : 0.0f,
                mass ? mass[i] : 1.0f);
}

size_t port_gen_fluid_block(const float* c, const float* s, float spacing, float m, size_t cap, float* pos3, float* mass) {
    sink_t k = {pos3, mass, cap, 0}; gen_fluid_block(&k, c, s, spacing, m); return k.n;
}
size_t port_gen_boundary_box(const float* c, const float* s, float spacing, float m, size_t cap, float* pos3, float* mass) {
    sink_t k = {pos3, mass, cap, 0}; gen_boundary_box(&k, c, s, spacing, m); return k.n;
}
size_t port_gen_fluid_drop(const float* c, float radius, float spacing, float m, size_t cap, float* pos3, float* mass) {
    sink_t k = {pos3, mass, cap, 0}; gen_fluid_drop(&k, c, radius, spacing, m); return k.n;
}
size_t port_gen_dam_break(const float* dam, const float* fluid, float spacing, float m, size_t cap, float* pos3, float* mass) {
    sink_t k = {pos3, mass, cap, 0}; gen_dam_break(&k, dam, fluid, spacing, m); return k.n;
}
void port_add_fluid_block(void* h, const float* c, const float* s, float spacing, float m) {
    gen_arg_t g; memcpy(g.a, c, 12); memcpy(g.b, s, 12); g.spacing = spacing; g.m = m; g.radius = 0;
    add_generated((engine_t*)h, fill_block, &g);
}
void port_add_boundary_box(void* h, const float* c, const float* s, float spacing, float m) {
    gen_arg_t g; memcpy(g.a, c, 12); memcpy(g.b, s, 12); g.spacing = spacing; g.m = m; g.radius = 0;
    add_generated((engine_t*)h, fill_box, &g);
}

size_t port_size(void* h) { return ((engine_t*)h)->n; }
size_t port_capacity(void* h) { return ((engine_t*)h)->cap; }
float port_time(void* h) { return ((engine_t*)h)->time; }
size_t port_step_count(void* h) { return ((engine_t*)h)->step_count; }
int port_is_initialized(void* h) { return ((engine_t*)h)->initialized; }
float port_cfl_timestep(void* h) { return compute_cfl_timestep((engine_t*)h); }
void port_update_neighbor_lists(void* h) { update_neighbor_lists((engine_t*)h); }

/* SPHEngine::step, sph_engine.cpp:93-144 */
void port_step(void* h, float dt) {
    engine_t* e = (engine_t*)h;
    if (!e->initialized || e->n == 0) return;
    double t0 = now_s();
    if (dt <= 0.0f) dt = compute_cfl_timestep(e);
    double a = now_s();
    update_neighbor_lists(e);
    double b = now_s();
    e->t_nbr += b - a;
    compute_densities(e);
    double c = now_s();
    e->t_rho += c - b;
    compute_pressures(e);
    double d = now_s();
    compute_forces(e);
    double f = now_s();
    e->t_force += f - d;
    integrate_leapfrog(e, dt);
    double g = now_s();
    e->t_int += g - f;
    apply_boundary_conditions(e);
    e->time += dt;
    e->step_count++;
    e->t_total += now_s() - t0;
}

/* SPHEngine::run_steps, sph_engine.cpp:146-150 */
void port_run_steps(void* h, size_t n, int adaptive) {
    engine_t* e = (engine_t*)h;
    for (size_t i = 0; i < n; ++i) port_step(h, adaptive ? 0.0f : e->prm.timestep);
}

double port_timed_steps(void* h, size_t n, float dt) {
    double t0 = now_s();
    for (size_t i = 0; i < n; ++i) port_step(h, dt);
    return now_s() - t0;
}

void port_get_state(void* h, float* pos3, float* vel3, float* mass, float* rho, float* P, float* acc3) {
    engine_t* e = (engine_t*)h;
    for (size_t i = 0; i < e->n; ++i) {
        if (pos3) { pos3[3 * i] = e->px[i]; pos3[3 * i + 1] = e->py[i]; pos3[3 * i + 2] = e->pz[i]; }
        if (vel3) { vel3[3 * i] = e->vx[i]; vel3[3 * i + 1] = e->vy[i]; vel3[3 * i + 2] = e->vz[i]; }
        if (mass) mass[i] = e->mass[i];
        if (rho) rho[i] = e->rho[i];
        if (P) P[i] = e->P[i];
        if (acc3) { acc3[3 * i] = e->ax[i]; acc3[3 * i + 1] = e->ay[i]; acc3[3 * i + 2] = e->az[i]; }
    }
}

void port_set_state(void* h, const float* pos3, const float* vel3) {
    engine_t* e = (engine_t*)h;
    for (size_t i = 0; i < e->n; ++i) {
        if (pos3) { e->px[i] = pos3[3 * i]; e->py[i] = pos3[3 * i + 1]; e->pz[i] = pos3[3 * i + 2]; }
        if (vel3) { e->vx[i] = vel3[3 * i]; e->vy[i] = vel3[3 * i + 1]; e->vz[i] = vel3[3 * i + 2]; }
    }
}

void port_get_keys(void* h, uint64_t* keys) {
    engine_t* e = (engine_t*)h;
    for (size_t i = 0; i < e->n; ++i) {
        int cx, cy, cz;
        grid_coords(e, e->px[i], e->py[i], e->pz[i], &cx, &cy, &cz);
        keys[i] = hash_cell(cx, cy, cz);
    }
}

void port_get_neighbor_counts(void* h, uint32_t* counts) {
    engine_t* e = (engine_t*)h;
    for (size_t i = 0; i < e->n; ++i) counts[i] = e->nbr_cnt[i];
}
size_t port_get_neighbor_list(void* h, size_t i, size_t cap, uint32_t* out) {
    engine_t* e = (engine_t*)h;
    const uint32_t* l = nbr_list(e, i);
    for (size_t k = 0; k < e->nbr_cnt[i] && k < cap; ++k) out[k] = l[k];
    return e->nbr_cnt[i];
}
/* Stable permutation by key of the last hash build (what std::stable_sort by key gives). */
void port_get_sorted_order(void* h, uint32_t* order) {
    engine_t* e = (engine_t*)h;
    memcpy(order, e->order, e->n * sizeof(uint32_t));
}
size_t port_hash_total_cells(void* h) { return ((engine_t*)h)->n_cells; }
size_t port_hash_max_per_cell(void* h) {
    engine_t* e = (engine_t*)h;
    size_t m = 0;
    for (size_t c = 0; c < e->n_cells; ++c)
        if ((size_t)(e->ustart[c + 1] - e->ustart[c]) > m) m = e->ustart[c + 1] - e->ustart[c];
    return m;
}

/* get_total_mass, sph_engine.cpp:187-190: serial fp32 accumulate of rho times h^3 */
float port_total_mass(void* h) {
    engine_t* e = (engine_t*)h;
    float s = 0.0f;
    for (size_t i = 0; i < e->n; ++i) s = s + e->rho[i];
    return s * (e->prm.smoothing_length * e->prm.smoothing_length * e->prm.smoothing_length);
}
/* get_total_energy, sph_engine.cpp:192-200 */
float port_total_energy(void* h) {
    engine_t* e = (engine_t*)h;
    float ke = 0.0f;
    for (size_t i = 0; i < e->n; ++i) {
        float s2 = e->vx[i] * e->vx[i] + e->vy[i] * e->vy[i] + e->vz[i] * e->vz[i];
        ke += 0.5f * e->mass[i] * s2;
    }
    return ke;
}
/* compute_conservation_errors, sph_engine.cpp:178-185 */
void port_conservation_errors(void* h, float* mass_err, float* energy_err) {
    engine_t* e = (engine_t*)h;
    float total = port_total_mass(h);
    float initial = (float)e->n * e->prm.particle_mass;
    *mass_err = fabsf(total - initial) / initial;
    *energy_err = 0.0f;
}

void port_get_stats(void* h, double* s7) {
    engine_t* e = (engine_t*)h;
    s7[0] = e->t_total; s7[1] = e->t_nbr; s7[2] = e->t_rho; s7[3] = e->t_force; s7[4] = e->t_int;
    s7[5] = (double)e->max_neighbors; s7[6] = (double)e->total_queries;
}

/* SPHEngine::get_densities() returns the capacity-length buffer (sph_engine.h:135; empty before initialize) */
size_t port_get_densities_raw(void* h, size_t cap, float* out) {
    engine_t* e = (engine_t*)h;
    if (!e->buffers_ready) return 0;
    for (size_t i = 0; i < e->cap && i < cap; ++i) out[i] = e->rho[i];
    return e->cap;
}

/* create_kernel(type, h), kernels.cpp:224-236, installed where the engine keeps its kernel (sph_engine.h:42).  Like the
 * reference, initialize() and set_smoothing_length() put the cubic spline back (sph_engine.cpp:23, 162). */
int port_set_kernel(void* h, int type) {
    engine_t* e = (engine_t*)h;
    if (type < 0 || type > 2) return -1;
    e->ktype = type;
    return 0;
}
float port_kernel_W(void* h, float x, float y, float z) { return kW((engine_t*)h, x, y, z); }
void port_kernel_gradW(void* h, float x, float y, float z, float* o) { kgradW((engine_t*)h, x, y, z, &o[0], &o[1], &o[2]); }
float port_kernel_lapW(void* h, float x, float y, float z) { return klapW((engine_t*)h, x, y, z); }
