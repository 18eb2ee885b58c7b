"""CPU: the numpy scene generators of the host layer are bit-identical to the reference's C++ generators."""
import hashlib
import json

import numpy as np
import pytest

from helpers import GOLDEN, assert_bits


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def scenes(pkg):
    from sph_b200 import scenes
    return scenes


def test_generator_hashes_match_reference(scenes):
    meta = json.loads((GOLDEN / "scalars.json").read_text())["generators"]
    for name, rec in meta.items():
        fam, dx = scenes.SCENES[name]
        if fam == "dam":
            pos, _ = scenes.create_dam_break_setup((0.4, 0.6, 0.8), (0.2, 0.4, 0.8), dx, 1.0)
        else:
            pos, _ = scenes.create_fluid_drop_setup((0.0, 0.5, 0.0), 0.1, dx, 1.0)
        assert pos.shape[0] == rec["n"], name
        assert sha(pos) == rec["pos_sha256"], name


def test_generators_match_oracle_port(scenes, po):
    for args in [((0, 0.3, 0), (0.4, 0.6, 0.8), 0.01), ((0.1, -0.2, 0.3), (0.37, 0.61, 0.83), 0.013)]:
        a, _ = scenes.create_fluid_block(*args, 1.0); b, _ = po.gen_fluid_block(*args, 1.0, kind="port")
        assert_bits(a, b, "fluid block")
        a, _ = scenes.create_boundary_box(*args, 1.0); b, _ = po.gen_boundary_box(*args, 1.0, kind="port")
        assert_bits(a, b, "boundary box")


def test_tame_params(scenes):
    p = scenes.tame_params(0.004, 0.008, scenes.DAM_BOUNDS)
    assert np.float32(p["neighbor_search_radius"]) == np.float32(2.0) * np.float32(0.008)
    assert abs(p["particle_mass"] - 1.5 * 1000 * 0.004 ** 3) < 1e-15
    assert p["xmin"] == -0.2 and p["zmax"] == 0.4
    pos, mass, prm, dt = scenes.make_scene("dam_break_13k")
    assert pos.shape == (13200, 3) and mass.shape == (13200,) and dt == prm["timestep"]


def test_dam_break_count_and_weak_scaling_spacing(scenes):
    """dam_break_count predicts the generator's yield exactly; the weak-scaling spacing gives N x the particles."""
    for dx in (0.02, 0.013, 0.01, 0.0077):
        assert scenes.dam_break_count(dx) == scenes.dam_break_scene(dx)[0].shape[0], dx
    n1 = scenes.dam_break_count(0.004)
    assert n1 == 1130000
    for world in (2, 4, 8):
        dx = scenes.dam_break_dx_for(world * n1, 0.004 / world ** (1.0 / 3.0))
        assert abs(scenes.dam_break_count(dx) - world * n1) <= 0.005 * world * n1
