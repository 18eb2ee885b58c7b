// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// Flat C driver over the UNMODIFIED reference engine (sph::SPHEngine compiled from
// /root/reference/src/{particle,spatial_hash,kernels,sph_engine}.cpp where they lie;
// see oracle/Makefile).  It exists so tests/golden/make_golden.py, the parity tests
// and bench.py's cpu_baseline / --impl reference legs can drive the reference through
// ctypes and read its per-stage state.  Built with -fno-access-control so the private
// per-step buffers (accelerations_, neighbor_lists_; reference sph_engine.h:62-65) and
// SpatialHash::hash_position (spatial_hash.h:20-27) can be observed without editing
// reference sources.
#include "sph_engine.h"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using sph::Particle;
using sph::SPHEngine;
using sph::SPHParameters;

namespace {

// params16 layout (shared with oracle/sph_oracle.c and include/sphb.h):
// 0 rest_density 1 gas_constant 2 viscosity 3 smoothing_length 4 particle_mass
// 5 timestep 6 gravity 7 damping 8 CFL_factor 9..14 xmin xmax ymin ymax zmin zmax
// 15 neighbor_search_radius
SPHParameters unpack(const float* p) {
    SPHParameters q;
    q.rest_density = p[0];
    q.gas_constant = p[1];
    q.viscosity = p[2];
    q.smoothing_length = p[3];
    q.particle_mass = p[4];
    q.timestep = p[5];
    q.gravity = p[6];
    q.damping = p[7];
    q.CFL_factor = p[8];
    q.bounds.xmin = p[9];
    q.bounds.xmax = p[10];
    q.bounds.ymin = p[11];
    q.bounds.ymax = p[12];
    q.bounds.zmin = p[13];
    q.bounds.zmax = p[14];
    q.neighbor_search_radius = p[15];
    return q;
}

void pack(const SPHParameters& q, float* p) {
    p[0] = q.rest_density;
    p[1] = q.gas_constant;
    p[2] = q.viscosity;
    p[3] = q.smoothing_length;
    p[4] = q.particle_mass;
    p[5] = q.timestep;
    p[6] = q.gravity;
    p[7] = q.damping;
    p[8] = q.CFL_factor;
    p[9] = q.bounds.xmin;
    p[10] = q.bounds.xmax;
    p[11] = q.bounds.ymin;
    p[12] = q.bounds.ymax;
    p[13] = q.bounds.zmin;
    p[14] = q.bounds.zmax;
    p[15] = q.neighbor_search_radius;
}

size_t copy_out(const std::vector<Particle>& v, size_t cap, float* pos3, float* vel3, float* mass) {
    size_t n = v.size() < cap ? v.size() : cap;
    for (size_t i = 0; i < n; ++i) {
        if (pos3) { pos3[3 * i] = v[i].position.x; pos3[3 * i + 1] = v[i].position.y; pos3[3 * i + 2] = v[i].position.z; }
        if (vel3) { vel3[3 * i] = v[i].velocity.x; vel3[3 * i + 1] = v[i].velocity.y; vel3[3 * i + 2] = v[i].velocity.z; }
        if (mass) mass[i] = v[i].mass;
    }
    return v.size();
}

}  // namespace

extern "C" {

void* ref_create(size_t max_particles) { return new SPHEngine(max_particles); }
void ref_destroy(void* e) { delete static_cast<SPHEngine*>(e); }

int ref_sizeof_particle() { return static_cast<int>(sizeof(Particle)); }
int ref_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ref_default_params(float* p16) { pack(SPHParameters{}, p16); }
void ref_initialize(void* e, const float* p16) { static_cast<SPHEngine*>(e)->initialize(unpack(p16)); }
void ref_set_parameters(void* e, const float* p16) { static_cast<SPHEngine*>(e)->set_parameters(unpack(p16)); }
void ref_get_parameters(void* e, float* p16) { pack(static_cast<SPHEngine*>(e)->get_parameters(), p16); }
void ref_set_smoothing_length(void* e, float h) { static_cast<SPHEngine*>(e)->set_smoothing_length(h); }
void ref_set_gravity(void* e, float g) { static_cast<SPHEngine*>(e)->set_gravity(g); }
void ref_set_viscosity(void* e, float mu) { static_cast<SPHEngine*>(e)->set_viscosity(mu); }
void ref_set_boundaries(void* e, float x0, float x1, float y0, float y1, float z0, float z1) {
    static_cast<SPHEngine*>(e)->set_boundaries(x0, x1, y0, y1, z0, z1);
}
void ref_initialize_dam_break(void* e) { static_cast<SPHEngine*>(e)->initialize_dam_break(); }
void ref_initialize_fluid_drop(void* e) { static_cast<SPHEngine*>(e)->initialize_fluid_drop(); }
void ref_initialize_granular_flow(void* e) { static_cast<SPHEngine*>(e)->initialize_granular_flow(); }
void ref_clear_particles(void* e) { static_cast<SPHEngine*>(e)->clear_particles(); }

// Arbitrary particle input through the reference's own add_particles.
void ref_add_particles(void* e, size_t n, const float* pos3, const float* vel3, const float* mass) {
    std::vector<Particle> v(n);
    for (size_t i = 0; i < n; ++i) {
        v[i].position = glm::vec3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
        if (vel3) v[i].velocity = glm::vec3(vel3[3 * i], vel3[3 * i + 1], vel3[3 * i + 2]);
        if (mass) v[i].mass = mass[i];
    }
    static_cast<SPHEngine*>(e)->add_particles(v);
}

// The reference's lattice generators (particle.cpp:166-229, sph_engine.cpp:489-514), exposed
// both as "add to engine" and as "just give me the list" (cap = room in the out arrays;
// return value = number generated).
size_t ref_gen_fluid_block(const float* c, const float* s, float spacing, float m, size_t cap, float* pos3, float* mass) {
    auto v = sph::create_fluid_block(glm::vec3(c[0], c[1], c[2]), glm::vec3(s[0], s[1], s[2]), spacing, m);
    return copy_out(v, cap, pos3, nullptr, mass);
}
size_t ref_gen_boundary_box(const float* c, const float* s, float spacing, float m, size_t cap, float* pos3, float* mass) {
    auto v = sph::create_boundary_box(glm::vec3(c[0], c[1], c[2]), glm::vec3(s[0], s[1], s[2]), spacing, m);
    return copy_out(v, cap, pos3, nullptr, mass);
}
size_t ref_gen_fluid_drop(const float* c, float radius, float spacing, float m, size_t cap, float* pos3, float* mass) {
    SPHParameters p;
    p.particle_mass = m;
    auto v = sph::utils::create_fluid_drop_setup(glm::vec3(c[0], c[1], c[2]), radius, spacing, p);
    return copy_out(v, cap, pos3, nullptr, mass);
}
size_t ref_gen_dam_break(const float* dam, const float* fluid, float spacing, float m, size_t cap, float* pos3, float* mass) {
    SPHParameters p;
    p.particle_mass = m;
    auto v = sph::utils::create_dam_break_setup(glm::vec3(dam[0], dam[1], dam[2]), glm::vec3(fluid[0], fluid[1], fluid[2]), spacing, p);
    return copy_out(v, cap, pos3, nullptr, mass);
}
void ref_add_fluid_block(void* e, const float* c, const float* s, float spacing, float m) {
    static_cast<SPHEngine*>(e)->add_particles(sph::create_fluid_block(glm::vec3(c[0], c[1], c[2]), glm::vec3(s[0], s[1], s[2]), spacing, m));
}
void ref_add_boundary_box(void* e, const float* c, const float* s, float spacing, float m) {
    static_cast<SPHEngine*>(e)->add_particles(sph::create_boundary_box(glm::vec3(c[0], c[1], c[2]), glm::vec3(s[0], s[1], s[2]), spacing, m));
}

size_t ref_size(void* e) { return static_cast<SPHEngine*>(e)->get_particles().size(); }
size_t ref_capacity(void* e) { return static_cast<SPHEngine*>(e)->get_particles().capacity(); }
float ref_time(void* e) { return static_cast<SPHEngine*>(e)->get_current_time(); }
size_t ref_step_count(void* e) { return static_cast<SPHEngine*>(e)->get_step_count(); }
int ref_is_initialized(void* e) { return static_cast<SPHEngine*>(e)->is_initialized() ? 1 : 0; }

void ref_step(void* e, float dt) { static_cast<SPHEngine*>(e)->step(dt); }
void ref_run_steps(void* e, size_t n, int adaptive) { static_cast<SPHEngine*>(e)->run_steps(n, adaptive != 0); }

// Wall-clock seconds for n steps (what bench.py times; excludes everything but step()).
double ref_timed_steps(void* e, size_t n, float dt) {
    auto t0 = std::chrono::high_resolution_clock::now();
    for (size_t i = 0; i < n; ++i) static_cast<SPHEngine*>(e)->step(dt);
    auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

float ref_cfl_timestep(void* e) { return static_cast<SPHEngine*>(e)->compute_cfl_timestep(); }

// State in insertion order.  Any pointer may be NULL.  rho/P come from the engine buffers
// (densities_/pressures_), acc from accelerations_.
void ref_get_state(void* e, float* pos3, float* vel3, float* mass, float* rho, float* P, float* acc3) {
    SPHEngine* s = static_cast<SPHEngine*>(e);
    const size_t n = s->particles_.size();
    for (size_t i = 0; i < n; ++i) {
        const Particle& p = s->particles_[i];
        if (pos3) { pos3[3 * i] = p.position.x; pos3[3 * i + 1] = p.position.y; pos3[3 * i + 2] = p.position.z; }
        if (vel3) { vel3[3 * i] = p.velocity.x; vel3[3 * i + 1] = p.velocity.y; vel3[3 * i + 2] = p.velocity.z; }
        if (mass) mass[i] = p.mass;
        if (rho) rho[i] = s->densities_[i];
        if (P) P[i] = s->pressures_[i];
        if (acc3) { acc3[3 * i] = s->accelerations_[i].x; acc3[3 * i + 1] = s->accelerations_[i].y; acc3[3 * i + 2] = s->accelerations_[i].z; }
    }
}

// Overwrite positions/velocities in place (teacher forcing).
void ref_set_state(void* e, const float* pos3, const float* vel3) {
    SPHEngine* s = static_cast<SPHEngine*>(e);
    const size_t n = s->particles_.size();
    for (size_t i = 0; i < n; ++i) {
        Particle& p = s->particles_[i];
        if (pos3) p.position = glm::vec3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
        if (vel3) p.velocity = glm::vec3(vel3[3 * i], vel3[3 * i + 1], vel3[3 * i + 2]);
    }
}

// Cell keys of the CURRENT positions, with the engine's own hash object
// (spatial_hash.h:20-36, private → -fno-access-control).
void ref_get_keys(void* e, uint64_t* keys) {
    SPHEngine* s = static_cast<SPHEngine*>(e);
    const size_t n = s->particles_.size();
    for (size_t i = 0; i < n; ++i) {
        glm::ivec3 c = s->spatial_hash_->get_grid_coords(s->particles_[i].position);
        keys[i] = s->spatial_hash_->hash_position(c.x, c.y, c.z);
    }
}

// Neighbour lists as built by the LAST step (i.e. for the pre-integration positions).
void ref_get_neighbor_counts(void* e, uint32_t* counts) {
    SPHEngine* s = static_cast<SPHEngine*>(e);
    const size_t n = s->particles_.size();
    for (size_t i = 0; i < n; ++i) counts[i] = static_cast<uint32_t>(s->neighbor_lists_[i].size());
}
size_t ref_get_neighbor_list(void* e, size_t i, size_t cap, uint32_t* out) {
    SPHEngine* s = static_cast<SPHEngine*>(e);
    const auto& l = s->neighbor_lists_[i];
    for (size_t k = 0; k < l.size() && k < cap; ++k) out[k] = static_cast<uint32_t>(l[k]);
    return l.size();
}

// Rebuild the neighbour lists for the current positions without stepping.
void ref_update_neighbor_lists(void* e) { static_cast<SPHEngine*>(e)->update_neighbor_lists(); }
size_t ref_hash_total_cells(void* e) { return static_cast<SPHEngine*>(e)->spatial_hash_->get_total_cells(); }
size_t ref_hash_max_per_cell(void* e) { return static_cast<SPHEngine*>(e)->spatial_hash_->get_max_particles_per_cell(); }

float ref_total_mass(void* e) { return static_cast<SPHEngine*>(e)->get_total_mass(); }
float ref_total_energy(void* e) { return static_cast<SPHEngine*>(e)->get_total_energy(); }
void ref_conservation_errors(void* e, float* mass_err, float* energy_err) {
    static_cast<SPHEngine*>(e)->compute_conservation_errors(*mass_err, *energy_err);
}

// stats7: total, neighbor, density, force, integration seconds, max_neighbors, total_queries
void ref_get_stats(void* e, double* stats7) {
    const auto& st = static_cast<SPHEngine*>(e)->get_performance_stats();
    stats7[0] = st.total_time;
    stats7[1] = st.neighbor_search_time;
    stats7[2] = st.density_computation_time;
    stats7[3] = st.force_computation_time;
    stats7[4] = st.integration_time;
    stats7[5] = static_cast<double>(st.max_neighbors);
    stats7[6] = static_cast<double>(st.total_neighbor_queries);
}
void ref_reset_stats(void* e) { static_cast<SPHEngine*>(e)->reset_performance_stats(); }

// Capacity-length engine buffers exactly as SPHEngine::get_densities()/get_pressures() return
// them (sph_engine.h:135-136).
size_t ref_get_densities_raw(void* e, size_t cap, float* out) {
    auto v = static_cast<SPHEngine*>(e)->get_densities();
    for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    return v.size();
}

// Smoothing-kernel known-answer probes (kernels.cpp:140-153 through the engine's own kernel object).
// create_kernel (kernels.cpp:224-236) installed into the engine's own kernel slot (sph_engine.h:42): the unmodified engine
// then runs its density / force passes through the reference's own Wendland C2 / Gaussian classes (virtual calls at
// sph_engine.cpp:209, 232, 236).  type: 0 cubic spline, 1 Wendland C2, 2 Gaussian (kernels.h KernelType order).
int ref_set_kernel(void* e, int type) {
    if (type < 0 || type > 2) return -1;
    auto* eng = static_cast<SPHEngine*>(e);
    const sph::KernelType kt = type == 0 ? sph::KernelType::CUBIC_SPLINE : (type == 1 ? sph::KernelType::WENDLAND_C2 : sph::KernelType::GAUSSIAN);
    eng->kernel_ = sph::create_kernel(kt, eng->params_.smoothing_length);
    return 0;
}
float ref_kernel_W(void* e, float x, float y, float z) { return static_cast<SPHEngine*>(e)->kernel_->W(glm::vec3(x, y, z)); }
void ref_kernel_gradW(void* e, float x, float y, float z, float* out3) {
    glm::vec3 g = static_cast<SPHEngine*>(e)->kernel_->gradW(glm::vec3(x, y, z));
    out3[0] = g.x; out3[1] = g.y; out3[2] = g.z;
}
float ref_kernel_lapW(void* e, float x, float y, float z) { return static_cast<SPHEngine*>(e)->kernel_->laplacianW(glm::vec3(x, y, z)); }

}  // extern "C"
