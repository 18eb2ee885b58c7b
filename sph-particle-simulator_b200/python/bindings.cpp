// pybind11 module `sph`: the Python face of the drop-in.  Same module name, class names, method
// names, keyword arguments and array shapes as the reference's python/bindings.cpp:8-166
// (Simulator == SPHEngine, ctor keyword max_particles, step(dt=0.0), run_steps(num_steps,
// adaptive_timestep=True), (N,3) float32 position/velocity copies, capacity-length density/pressure
// arrays, create_fluid_block/create_boundary_box(center, size, spacing, mass=1.0), __version__).
// Differences, all additive: vec3 fields of Particle are readable/writable as 3-tuples (the reference
// registers them without a caster, so touching them raises TypeError there), and Simulator gains
// set_math_mode / get_accelerations / compute_cfl_timestep.
#include <pybind11/numpy.h>
#include <cstring>

#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <stdexcept>

#include "sph_host.h"

namespace py = pybind11;
using sph::Particle;
using sph::ParticleSystem;
using sph::SPHEngine;
using sph::SPHParameters;

namespace {

py::tuple to_tuple(const glm::vec3& v) { return py::make_tuple(v.x, v.y, v.z); }

glm::vec3 to_vec3(const py::handle& h) {
    py::sequence s = py::reinterpret_borrow<py::sequence>(h);
    if (py::len(s) != 3) throw std::invalid_argument("expected a 3-vector");
    return glm::vec3(s[0].cast<float>(), s[1].cast<float>(), s[2].cast<float>());
}

glm::vec3 arr3(const py::array_t<float>& a, const char* what) {
    if (a.size() != 3) throw std::invalid_argument(std::string(what) + " must be a 3D vector");
    return glm::vec3(a.at(0), a.at(1), a.at(2));
}

py::array_t<float> vec3_array(const std::vector<glm::vec3>& v) {
    py::array_t<float> out({static_cast<py::ssize_t>(v.size()), static_cast<py::ssize_t>(3)});
    if (!v.empty()) std::memcpy(out.mutable_data(), v.data(), v.size() * sizeof(glm::vec3));
    return out;
}

py::array_t<float> float_array(const std::vector<float>& v) {
    py::array_t<float> out(static_cast<py::ssize_t>(v.size()));
    if (!v.empty()) std::memcpy(out.mutable_data(), v.data(), v.size() * sizeof(float));
    return out;
}

#define VEC3_PROPERTY(cls, name)                                              \
    .def_property(#name, [](const cls& p) { return to_tuple(p.name); },        \
                  [](cls& p, const py::object& v) { p.name = to_vec3(v); })

}  // namespace

PYBIND11_MODULE(sph, m) {
    m.doc() = "SPH Particle Simulator - Python bindings (B200-native hot path behind the reference API)";

    py::class_<SPHParameters>(m, "SPHParameters")
        .def(py::init<>())
        .def_readwrite("rest_density", &SPHParameters::rest_density)
        .def_readwrite("gas_constant", &SPHParameters::gas_constant)
        .def_readwrite("viscosity", &SPHParameters::viscosity)
        .def_readwrite("smoothing_length", &SPHParameters::smoothing_length)
        .def_readwrite("particle_mass", &SPHParameters::particle_mass)
        .def_readwrite("timestep", &SPHParameters::timestep)
        .def_readwrite("gravity", &SPHParameters::gravity)
        .def_readwrite("damping", &SPHParameters::damping);

    py::enum_<sph::ParticleType>(m, "ParticleType")
        .value("FLUID", sph::ParticleType::FLUID)
        .value("BOUNDARY", sph::ParticleType::BOUNDARY)
        .value("SOLID", sph::ParticleType::SOLID);

    py::class_<Particle>(m, "Particle")
        .def(py::init<>())
        .def(py::init([](const py::object& pos, float mass, sph::ParticleType type) { return Particle(to_vec3(pos), mass, type); }),
             py::arg("position"), py::arg("mass") = 1.0f, py::arg("type") = sph::ParticleType::FLUID)
        VEC3_PROPERTY(Particle, position)
        VEC3_PROPERTY(Particle, velocity)
        VEC3_PROPERTY(Particle, acceleration)
        .def_readwrite("density", &Particle::density)
        .def_readwrite("pressure", &Particle::pressure)
        .def_readwrite("mass", &Particle::mass)
        .def_readwrite("type", &Particle::type)
        .def_readwrite("temperature", &Particle::temperature)
        .def_readwrite("viscosity", &Particle::viscosity)
        VEC3_PROPERTY(Particle, color)
        .def_readonly("id", &Particle::id);

    py::class_<ParticleSystem>(m, "ParticleSystem")
        .def("size", &ParticleSystem::size)
        .def("capacity", &ParticleSystem::capacity)
        .def("empty", &ParticleSystem::empty)
        .def("get_positions", [](const ParticleSystem& ps) { return vec3_array(ps.get_positions()); })
        .def("get_velocities", [](const ParticleSystem& ps) { return vec3_array(ps.get_velocities()); })
        .def("get_densities", [](const ParticleSystem& ps) { return float_array(ps.get_densities()); })
        .def("get_pressures", [](const ParticleSystem& ps) { return float_array(ps.get_pressures()); });

    py::class_<SPHEngine::PerformanceStats>(m, "PerformanceStats")
        .def_readonly("total_time", &SPHEngine::PerformanceStats::total_time)
        .def_readonly("neighbor_search_time", &SPHEngine::PerformanceStats::neighbor_search_time)
        .def_readonly("density_computation_time", &SPHEngine::PerformanceStats::density_computation_time)
        .def_readonly("force_computation_time", &SPHEngine::PerformanceStats::force_computation_time)
        .def_readonly("integration_time", &SPHEngine::PerformanceStats::integration_time)
        .def_readonly("max_neighbors", &SPHEngine::PerformanceStats::max_neighbors)
        .def_readonly("total_neighbor_queries", &SPHEngine::PerformanceStats::total_neighbor_queries);

    py::class_<SPHEngine>(m, "Simulator")
        .def(py::init<size_t>(), py::arg("max_particles") = 1000000)
        .def("initialize", &SPHEngine::initialize, py::arg("params"))
        .def("initialize_dam_break", &SPHEngine::initialize_dam_break)
        .def("initialize_fluid_drop", &SPHEngine::initialize_fluid_drop)
        .def("initialize_granular_flow", &SPHEngine::initialize_granular_flow)
        .def("add_particles", &SPHEngine::add_particles)
        .def("clear_particles", &SPHEngine::clear_particles)
        .def("step", &SPHEngine::step, py::arg("dt") = 0.0f)
        .def("run_steps", &SPHEngine::run_steps, py::arg("num_steps"), py::arg("adaptive_timestep") = true)
        .def("get_particles", &SPHEngine::get_particles, py::return_value_policy::reference)
        .def("get_parameters", &SPHEngine::get_parameters)
        .def("get_current_time", &SPHEngine::get_current_time)
        .def("get_step_count", &SPHEngine::get_step_count)
        .def("set_parameters", &SPHEngine::set_parameters)
        .def("set_gravity", &SPHEngine::set_gravity)
        .def("set_viscosity", &SPHEngine::set_viscosity)
        .def("set_smoothing_length", &SPHEngine::set_smoothing_length)
        .def("set_boundaries", &SPHEngine::set_boundaries)
        .def("get_performance_stats", &SPHEngine::get_performance_stats)
        .def("reset_performance_stats", &SPHEngine::reset_performance_stats)
        .def("compute_conservation_errors",
             [](SPHEngine& e) {
                 float mass_error = 0.0f, energy_error = 0.0f;
                 e.compute_conservation_errors(mass_error, energy_error);
                 return py::make_tuple(mass_error, energy_error);
             })
        .def("get_total_mass", &SPHEngine::get_total_mass)
        .def("get_total_energy", &SPHEngine::get_total_energy)
        .def("validate_simulation", &SPHEngine::validate_simulation)
        .def("is_initialized", &SPHEngine::is_initialized)
        .def("get_positions", [](const SPHEngine& e) { return vec3_array(e.get_positions()); })
        .def("get_velocities", [](const SPHEngine& e) { return vec3_array(e.get_velocities()); })
        .def("get_densities", [](const SPHEngine& e) { return float_array(e.get_densities()); })
        .def("get_pressures", [](const SPHEngine& e) { return float_array(e.get_pressures()); })
        // additions
        .def("set_math_mode", &SPHEngine::set_math_mode, py::arg("mode"))
        .def("set_kernel_type", &SPHEngine::set_kernel_type, py::arg("type"),
             "0 cubic spline (default), 1 Wendland C2, 2 Gaussian (reference KernelType order)")
        .def("set_async", &SPHEngine::set_async, py::arg("on"))
        .def("get_accelerations", [](const SPHEngine& e) { return vec3_array(e.get_accelerations()); })
        .def("compute_cfl_timestep", &SPHEngine::compute_cfl_timestep)
        .def("device_count", &SPHEngine::device_count, "GPUs this engine owns (environment SPHB_DEVICES=0,1,... at construction; default 1)")
        .def("get_instance_data",
             [](const SPHEngine& e) {
                 const std::vector<float> v = e.get_instance_data();
                 py::array_t<float> a({static_cast<py::ssize_t>(v.size() / 9), static_cast<py::ssize_t>(9)});
                 std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(float));
                 return a;
             },
             "renderer instance records (N, 9): position, velocity, colour — one kernel instead of the per-frame AoS walk")
        .def("get_report_diagnostics",
             [](const SPHEngine& e) {
                 const SPHEngine::ReportDiagnostics d = e.get_report_diagnostics();
                 py::dict out;
                 out["mass_error"] = d.mass_error;
                 out["kinetic_energy"] = d.kinetic_energy;
                 out["average_density"] = d.average_density;
                 out["max_velocity"] = d.max_velocity;
                 out["total_mass"] = d.total_mass;
                 return out;
             },
             "mass error, kinetic energy, mean density over the capacity-long buffer and max |v| from one device reduction");

    m.def("create_fluid_block",
          [](const py::array_t<float>& center, const py::array_t<float>& size, float spacing, float mass) {
              return sph::create_fluid_block(arr3(center, "center and size"), arr3(size, "center and size"), spacing, mass);
          },
          py::arg("center"), py::arg("size"), py::arg("spacing"), py::arg("mass") = 1.0f);
    m.def("create_boundary_box",
          [](const py::array_t<float>& center, const py::array_t<float>& size, float spacing, float mass) {
              return sph::create_boundary_box(arr3(center, "center and size"), arr3(size, "center and size"), spacing, mass);
          },
          py::arg("center"), py::arg("size"), py::arg("spacing"), py::arg("mass") = 1.0f);

    m.attr("__version__") = "1.0.0";
}
