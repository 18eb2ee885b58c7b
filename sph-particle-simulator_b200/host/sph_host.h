// Host-side shell with the reference's public C++ surface: sph::Particle, sph::ParticleSystem,
// sph::SPHParameters, sph::SPHEngine and the lattice generators — so code written against the
// reference's headers (src/particle.h, src/sph_engine.h) compiles against this one unchanged.
//
// What differs is where the work happens: SPHEngine owns an sphb_ctx (include/sphb.h) and forwards the
// per-step hot path to the GPU through that C ABI.  The host array-of-structs stays the
// authoritative container for everything the reference does on the host (scene building, capacity
// capping, ids) and is synchronised lazily: uploaded before the first step after a host-side change,
// refreshed from the device only when a caller actually looks at it.
#pragma once

#include <glm/glm.hpp>   // real GLM if installed, else include/compat/glm/glm.hpp

#include <cstddef>
#include <cstdint>
#include <vector>

struct sphb_ctx;
struct sphb_multi;

namespace sph {

// reference src/particle.h:10-14
enum class ParticleType { FLUID, BOUNDARY, SOLID };

// reference src/particle.h:17-49 — identical field order and defaults (76 bytes); the C ABI's strided
// upload/download reads and writes this layout in place.
struct Particle {
    glm::vec3 position{0.0f};
    glm::vec3 velocity{0.0f};
    glm::vec3 acceleration{0.0f};
    float density = 0.0f;
    float pressure = 0.0f;
    float mass = 1.0f;
    ParticleType type = ParticleType::FLUID;
    float temperature = 293.15f;
    float viscosity = 0.001f;
    glm::vec3 color{0.0f, 0.5f, 1.0f};
    int id = -1;

    Particle() = default;
    explicit Particle(const glm::vec3& pos, float m = 1.0f, ParticleType t = ParticleType::FLUID)
        : position(pos), mass(m), type(t) {}
};
static_assert(sizeof(Particle) == 76, "Particle must keep the reference's 76-byte layout");

// reference src/particle.h:52-100
class ParticleSystem {
public:
    explicit ParticleSystem(size_t capacity = 1000000);

    void reserve(size_t capacity);
    void resize(size_t size) { items_.resize(size); }
    void clear() { items_.clear(); }

    Particle& operator[](size_t i) { return items_[i]; }
    const Particle& operator[](size_t i) const { return items_[i]; }
    size_t size() const { return items_.size(); }
    size_t capacity() const { return capacity_; }
    bool empty() const { return items_.empty(); }

    auto begin() { return items_.begin(); }
    auto end() { return items_.end(); }
    auto begin() const { return items_.cbegin(); }
    auto end() const { return items_.cend(); }

    void add_particle(const Particle& p);
    void add_particles(const std::vector<Particle>& ps);
    void remove_particle(size_t index);
    void remove_particles(const std::vector<size_t>& indices);

    std::vector<glm::vec3> get_positions() const;
    std::vector<glm::vec3> get_velocities() const;
    std::vector<float> get_densities() const;
    std::vector<float> get_pressures() const;

    void set_mass(float mass);
    void set_viscosity(float viscosity);
    void set_temperature(float temperature);

    void apply_boundary_conditions(float xmin, float xmax, float ymin, float ymax, float zmin, float zmax);

    Particle* data() { return items_.data(); }
    const Particle* data() const { return items_.data(); }

private:
    std::vector<Particle> items_;
    size_t capacity_;
};

glm::vec3 generate_random_position(float xmin, float xmax, float ymin, float ymax, float zmin, float zmax);
std::vector<Particle> create_fluid_block(const glm::vec3& center, const glm::vec3& size, float spacing, float mass = 1.0f);
std::vector<Particle> create_boundary_box(const glm::vec3& center, const glm::vec3& size, float spacing, float mass = 1.0f);

// reference src/sph_engine.h:14-34
struct SPHParameters {
    float rest_density = 1000.0f;
    float gas_constant = 2000.0f;
    float viscosity = 0.001f;
    float smoothing_length = 0.02f;
    float particle_mass = 0.001f;
    float timestep = 0.001f;
    float gravity = -9.81f;
    float damping = 0.99f;
    float CFL_factor = 0.4f;
    struct Boundaries {
        float xmin = -1.0f, xmax = 1.0f;
        float ymin = -1.0f, ymax = 1.0f;
        float zmin = -1.0f, zmax = 1.0f;
    } bounds;
    float neighbor_search_radius = 0.04f;
};

// reference src/sph_engine.h:37-141
class SPHEngine {
public:
    // reference sph_engine.h:51-59 (private there although a public getter returns it; public here so
    // bindings can name it)
    struct PerformanceStats {
        double total_time = 0.0;
        double neighbor_search_time = 0.0;
        double density_computation_time = 0.0;
        double force_computation_time = 0.0;
        double integration_time = 0.0;
        size_t max_neighbors = 0;
        size_t total_neighbor_queries = 0;
    };

    explicit SPHEngine(size_t max_particles = 1000000);
    ~SPHEngine();
    SPHEngine(const SPHEngine&) = delete;
    SPHEngine& operator=(const SPHEngine&) = delete;

    void initialize(const SPHParameters& params);
    void initialize_dam_break();
    void initialize_fluid_drop();
    void initialize_granular_flow();

    void add_particles(const std::vector<Particle>& particles);
    void clear_particles();

    void step(float dt = 0.0f);   // dt <= 0: adaptive CFL timestep
    void run_steps(size_t num_steps, bool adaptive_timestep = true);

    const ParticleSystem& get_particles() const;
    const SPHParameters& get_parameters() const { return params_; }
    float get_current_time() const;
    size_t get_step_count() const { return step_count_; }

    void set_parameters(const SPHParameters& params);
    void set_gravity(float gravity) { params_.gravity = gravity; }
    void set_viscosity(float viscosity) { params_.viscosity = viscosity; }
    void set_smoothing_length(float h);
    void set_boundaries(float xmin, float xmax, float ymin, float ymax, float zmin, float zmax);

    const PerformanceStats& get_performance_stats() const;
    void reset_performance_stats();

    void compute_conservation_errors(float& mass_error, float& energy_error) const;
    // benchmarks/performance_test.cpp:157 passes doubles (which does not compile against the reference's own
    // float& signature); this overload lets that driver build unmodified
    void compute_conservation_errors(double& mass_error, double& energy_error) const;
    float get_total_mass() const;
    float get_total_energy() const;

    std::vector<glm::vec3> get_positions() const;
    std::vector<glm::vec3> get_velocities() const;
    std::vector<float> get_densities() const;   // capacity-length, like the reference's buffer
    std::vector<float> get_pressures() const;

    void validate_simulation() const;
    bool is_initialized() const { return initialized_; }

    // ---- additions (not in the reference) --------------------------------------------------------
    // 0 = bit-exact reference arithmetic, 1 = fast (default); see include/sphb.h SPHB_OPT_MATH_MODE
    void set_math_mode(int mode);
    // the smoothing kernel, reference KernelType order (kernels.h:94-98): 0 cubic spline (what the reference engine
    // hard-wires), 1 Wendland C2, 2 Gaussian — the other classes behind create_kernel; see SPHB_OPT_KERNEL_TYPE
    void set_kernel_type(int type);
    // accelerations of the last step in insertion order (the reference keeps them private)
    std::vector<glm::vec3> get_accelerations() const;
    // the CFL timestep the next adaptive step would take (reference: private compute_cfl_timestep)
    float compute_cfl_timestep() const;
    // false (default): step() returns when the GPU finished it, like the reference; true: enqueue and return
    void set_async(bool on);
    // What the reference's drivers derive from full-array getters every report (examples/dam_break.cpp:132-166:
    // compute_conservation_errors, get_total_energy, the mean of get_densities() over the CAPACITY-long buffer (quirk
    // Q13), the maximum of |get_velocities()|), from ONE device reduction — no array leaves the GPU.
    struct ReportDiagnostics {
        float mass_error;        // |sum(rho) h^3 - N m| / (N m)      (sph_engine.cpp:178-190)
        float kinetic_energy;    // sum 1/2 m |v|^2                   (sph_engine.cpp:192-200)
        float average_density;   // sum(rho) / capacity               (dam_break.cpp:146-151)
        float max_velocity;      // max |v|                           (dam_break.cpp:139-144)
        float total_mass;        // sum(rho) h^3
    };
    ReportDiagnostics get_report_diagnostics() const;
    // The renderer's instance buffer (Renderer::update_particle_data, src/renderer.cpp:279-312: position, velocity,
    // Particle::color — 9 floats per particle, insertion order) written by one kernel.  dst_on_device: dst is device
    // memory (a CUDA-mapped vertex buffer); otherwise host memory of size() * 9 floats.
    void export_instance_data(float* dst, bool dst_on_device = false) const;
    std::vector<float> get_instance_data() const;
    sphb_ctx* native_handle() const { return ctx_; }          // NULL when the engine owns several devices
    sphb_multi* native_multi_handle() const { return multi_; }
    int device_count() const;

private:
    void push_to_device() const;     // host AoS → device, if the host side changed
    void pull_from_device() const;   // device → host AoS, if the device advanced
    void tune_grid(float spacing);   // device grid refine from the generator's lattice spacing (tuning only)
    void push_params() const;
    [[noreturn]] void die(const char* what) const;

    mutable ParticleSystem particles_;
    SPHParameters params_;
    sphb_ctx* ctx_ = nullptr;       // one device (default)
    sphb_multi* multi_ = nullptr;   // several devices (environment SPHB_DEVICES=0,1,...)
    size_t step_count_ = 0;
    bool initialized_ = false;
    mutable bool host_changed_ = true;   // device does not hold the host particles yet
    mutable bool device_ahead_ = false;  // device state is newer than the host AoS
    mutable bool colors_on_device_ = false;
    mutable PerformanceStats perf_;
    int math_mode_ = 1;
    bool async_ = false;
};

// reference src/sph_engine.h:144-148 (never referenced by the engine; kept for source compatibility)
enum class IntegrationMethod { EULER, VERLET, LEAPFROG };

namespace utils {
std::vector<Particle> create_dam_break_setup(const glm::vec3& dam_size, const glm::vec3& fluid_size, float spacing,
                                             const SPHParameters& params);
std::vector<Particle> create_fluid_drop_setup(const glm::vec3& center, float radius, float spacing,
                                              const SPHParameters& params);
std::vector<Particle> create_granular_flow_setup(const glm::vec3& pile_size, const glm::vec3& domain_size, float spacing,
                                                 const SPHParameters& params);
bool validate_particle_setup(const std::vector<Particle>& particles);
float compute_average_neighbors(const SPHEngine& engine);
}  // namespace utils

}  // namespace sph
