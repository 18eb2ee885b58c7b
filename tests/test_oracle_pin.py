"""CPU, build container only: pin the plain-C oracle against the UNMODIFIED reference run live
(oracle/_ref/liboracle_strict.so).  Skipped where the compiled reference is not present."""
import numpy as np
import pytest

from helpers import assert_bits


@pytest.fixture(scope="module")
def strict(po):
    if not po.available("strict"):
        pytest.skip("oracle/_ref/liboracle_strict.so not built (needs /root/reference)")
    return po


@pytest.mark.parametrize("scene,steps", [("initialize_dam_break", 2), ("initialize_fluid_drop", 3), ("initialize_granular_flow", 2)])
def test_default_scenes_bit_exact(strict, scene, steps):
    po = strict
    a, b = po.Engine("strict", 120000), po.Engine("port", 120000)
    for e in (a, b):
        getattr(e, scene)()
    assert a.size == b.size
    for k in range(steps):
        assert_bits(a.keys(), b.keys(), f"{scene} keys before step {k}")
        a.step(0.001); b.step(0.001)
        sa, sb = a.state(), b.state()
        assert_bits(a.neighbor_counts(), b.neighbor_counts(), f"{scene} counts step {k}")
        for f in ("rho", "P", "acc", "pos", "vel"):
            assert_bits(sa[f], sb[f], f"{scene} step {k} {f}")
    assert a.total_mass() == b.total_mass() and a.total_energy() == b.total_energy()
    assert a.stats()["max_neighbors"] == b.stats()["max_neighbors"]
    a.close(); b.close()


def test_tame_dam_break_adaptive_bit_exact(strict, graft):
    po = strict
    graft.load_package()
    from sph_b200 import scenes
    pos, mass, prm, dt = scenes.dam_break_scene(0.01)
    a, b = po.Engine("strict", pos.shape[0]), po.Engine("port", pos.shape[0])
    for e in (a, b):
        e.initialize(prm); e.add_particles(pos, None, mass)
    for k in range(4):
        assert np.float32(a.cfl_timestep()) == np.float32(b.cfl_timestep())
        a.step(0.0); b.step(0.0)
        assert np.float32(a.time) == np.float32(b.time)
    sa, sb = a.state(), b.state()
    for f in ("rho", "P", "acc", "pos", "vel"):
        assert_bits(sa[f], sb[f], f"adaptive {f}")
    a.close(); b.close()


def test_neighbor_list_order(strict):
    """Lists themselves (not just counts) in the reference's order: dx, dy, dz ascending, ascending id in a cell."""
    po = strict
    a, b = po.Engine("strict", 10000), po.Engine("port", 10000)
    for e in (a, b):
        e.initialize_fluid_drop(); e.update_neighbor_lists()
    for i in (0, 1, 17, 4000, 8143):
        assert np.array_equal(a.neighbor_list(i), b.neighbor_list(i))
    a.close(); b.close()


def test_generators(strict):
    po = strict
    for args in [((0, 0.3, 0), (0.4, 0.6, 0.8), 0.01), ((0.1, -0.2, 0.3), (0.37, 0.61, 0.83), 0.013)]:
        for gen in (po.gen_fluid_block, po.gen_boundary_box):
            pa, ma = gen(*args, 0.5, kind="strict"); pb, mb = gen(*args, 0.5, kind="port")
            assert_bits(pa, pb, gen.__name__); assert_bits(ma, mb, gen.__name__)
    pa, _ = po.gen_fluid_drop((0, 0.5, 0), 0.1, 0.004, 1.0, kind="strict"); pb, _ = po.gen_fluid_drop((0, 0.5, 0), 0.1, 0.004, 1.0, kind="port")
    assert_bits(pa, pb, "drop")


@pytest.mark.parametrize("kernel_type", [1, 2])
def test_other_kernel_classes_bit_exact(strict, kernel_type):
    """the reference's Wendland C2 / Gaussian classes (kernels.cpp:166-222), installed in the engine's kernel
    slot by create_kernel, against the port's restatement: W / gradW / laplacianW at random separations and at the special
    points, then two engine steps; initialize() and set_smoothing_length() put the cubic spline back in both."""
    po = strict
    a, b = po.Engine("strict", 9000), po.Engine("port", 9000)
    for e in (a, b):
        e.initialize_fluid_drop()
        e.set_kernel(kernel_type)
    h = float(a.get_parameters()["smoothing_length"])
    rng = np.random.default_rng(5)
    rs = (rng.uniform(-1, 1, (3000, 3)) * 2.2 * h).astype(np.float32).tolist() + [[0, 0, 0], [2 * h, 0, 0], [h, 0, 0], [1e-7, 0, 0], [0, 3 * h, 0]]
    assert_bits(np.array([a.kernel_W(r) for r in rs], np.float32), np.array([b.kernel_W(r) for r in rs], np.float32), "W")
    assert_bits(np.array([a.kernel_gradW(r) for r in rs], np.float32), np.array([b.kernel_gradW(r) for r in rs], np.float32), "gradW")
    assert_bits(np.array([a.kernel_lapW(r) for r in rs], np.float32), np.array([b.kernel_lapW(r) for r in rs], np.float32), "lapW")
    for k in range(2):
        a.step(0.0005); b.step(0.0005)
        sa, sb = a.state(), b.state()
        for f in ("rho", "P", "acc", "pos", "vel"):
            assert_bits(sa[f], sb[f], f"kernel {kernel_type} step {k} {f}")
    for e in (a, b):
        e.set_smoothing_length(h)
    assert np.float32(a.kernel_W((0, 0, 0))) == np.float32(b.kernel_W((0, 0, 0)))
    sigma = np.float32(1.0) / (np.float32(np.pi) * np.float32(h) * np.float32(h) * np.float32(h))
    assert np.float32(b.kernel_W((0, 0, 0))) == sigma * np.float32(2.0 / 3.0)       # the cubic spline again (sph_engine.cpp:162)
    a.close(); b.close()
