"""Shared fixtures.  `-m "not gpu"` runs here (no GPU): the oracle, the host logic through the CPU
emulation of the device code, the ABI surface.  `-m gpu` runs on a B200: parity through the C ABI."""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# small versions of every named scene (SURVEY.md §4 example projects + BASELINE configs)
def small_scene_kwargs():
    from pyrite_b200 import scenes

    return {
        "cornell": dict(width=64, height=64, spp=4),
        "spheres": dict(width=64, height=48, spp=4),
        "diamonds": dict(width=64, height=48, spp=4),
        "textures": dict(width=64, height=48, spp=4),
        "rgb_emission": dict(width=64, height=48, spp=4),
        "snowflake": dict(width=64, height=48, spp=4),
        "fractals": dict(width=48, height=32, spp=2),
        # Julia regular / bicomplex, Mandelbulb with constant, bounds.sphere, image.filter
        "fractal_variants": dict(width=48, height=24, spp=2),
        "dragon": dict(width=64, height=48, spp=4, mesh=scenes.dragon_mesh(300, 24)),
        # limits: portrait film (the vertical AspectRatio branch, film.rs:203-246), maximum spectrum / light samples
        "edge_portrait": dict(_scene="cornell", width=40, height=72, spp=4, spectrum_samples=16, light_samples=8, bounces=2),
        # bidirectional integrator
        "bd_cornell": dict(_scene="cornell", integrator="bidirectional", width=48, height=48, spp=4),
        "bd_cornell_fractal": dict(_scene="cornell", integrator="bidirectional", fractal=True, width=48, height=48, spp=2),
        "bd_glass_dragon": dict(_scene="dragon", integrator="bidirectional", glass=True, bounces=12, light_bounces=20, width=48, height=32, spp=4,
                                mesh=scenes.dragon_mesh(300, 24)),
        "bd_diamonds": dict(_scene="diamonds", integrator="bidirectional", bounces=16, width=48, height=32, spp=4),
        "bd_spheres": dict(_scene="spheres", integrator="bidirectional", width=48, height=32, spp=4),
        "bd_c5": dict(_scene="bdpt_cornell_dragon", width=48, height=32, spp=2, mesh=scenes.dragon_mesh(300, 24)),
    }


SCENE_NAMES = ["cornell", "spheres", "diamonds", "textures", "rgb_emission", "snowflake", "fractals", "fractal_variants", "dragon", "edge_portrait", "lua_orbs"]
MARCHED_SCENES = ["fractals", "fractal_variants", "lua_orbs"]  # sphere-traced shapes: tolerance-level parity on the device
BIDIR_NAMES = ["bd_cornell", "bd_cornell_fractal", "bd_glass_dragon", "bd_diamonds", "bd_spheres", "bd_c5"]
MESH_SCENES = ["cornell", "diamonds", "textures", "snowflake", "dragon", "spheres", "rgb_emission"]

_ir_cache = {}


def scene_ir(name, **override):
    from pyrite_b200 import project, scenes

    key = (name, tuple(sorted((k, v) for k, v in override.items() if not hasattr(v, "position"))))
    if name == "lua_orbs":  # a project.lua through the Lua loader: directional light, bounds.sphere, clamp, curve spectra
        from pyrite_b200 import lua_project

        if key not in _ir_cache:
            _ir_cache[key] = lua_project.load_project_ir(ROOT / "tests" / "golden" / "scenes" / "orbs.lua")
        return _ir_cache[key]
    if key not in _ir_cache:
        kw = dict(small_scene_kwargs()[name])
        kw.update(override)
        fn = scenes.SCENES[kw.pop("_scene", name)]
        _ir_cache[key] = project.serialize_project(fn(**kw))
    return _ir_cache[key]


@pytest.fixture(scope="session")
def oracle_factory():
    from oracle_lib import Oracle

    made = {}

    def make(name, **override):
        key = (name, tuple(sorted(override.items())))
        if key not in made:
            made[key] = Oracle(scene_ir(name, **override))
        return made[key]

    return make


@pytest.fixture(scope="session")
def emu_factory():
    from emu_lib import Emu
    from oracle_lib import Oracle

    def make(name, **override):
        ir = scene_ir(name, **override)
        o = Oracle(ir)
        return Emu(ir, (o.info.height, o.info.width, o.info.bins)), o

    return make


@pytest.fixture(scope="session")
def gpu_renderer_factory():
    from pyrite_b200 import api

    live = []

    def make(name, **override):
        r = api.Renderer(0)
        r.load(scene_ir(name, **override))
        live.append(r)
        return r

    yield make
    for r in live:
        r.close()


os.environ.setdefault("OMP_NUM_THREADS", "4")
