"""Named scenes: the reference's example projects (pyrite/test/*/*.lua) and the BASELINE.json
configs C1-C5, written with the Python mirror of the Lua DSL (pyrite_b200.project).

Assets come from tests/golden/ (fixtures generated from the reference's own test assets by
tests/golden/make_fixtures.py).  `dragon.obj` is absent from the reference
(.MISSING_LARGE_BLOBS), so `dragon_mesh()` builds a deterministic 871,200-triangle stand-in
(SURVEY.md §8d C2).
"""
from __future__ import annotations

import json
from functools import lru_cache
from pathlib import Path

import numpy as np

from .project import (Image, Mesh, blackbody, bounds, camera, fresnel, light, light_source, material, mix,
                      quaternion_julia, ray_marched, renderer, rgb, shape, spectrum, texture, transform, vector)

ASSETS = Path(__file__).resolve().parent.parent / "tests" / "golden"


@lru_cache(maxsize=None)
def _mesh(name: str) -> Mesh:
    return Mesh.load(ASSETS / "meshes" / f"{name}.npz")


@lru_cache(maxsize=None)
def _spectra():
    return json.loads((ASSETS / "spectra.json").read_text())


@lru_cache(maxsize=None)
def _image(name: str) -> Image:
    return Image.open(ASSETS / "textures" / name)


# --------------------------------------------------------------------------- synthetic meshes
@lru_cache(maxsize=4)
def dragon_mesh(n_along: int = 1980, n_around: int = 220, seed: int = 0xD8A60) -> Mesh:
    """Closed, noise-displaced (2,3) torus-knot tube: 2 * n_along * n_around triangles
    (defaults: 871,200 near-isotropic triangles, edge lengths about 0.05 x 0.07 - a scanned mesh
    such as the Stanford dragon has well-shaped triangles, so the stand-in avoids slivers),
    object name `dragon`, smooth vertex normals, no UVs.  Model-space
    bounding box about 8 x 19 x 14 standing on z = 0 (so that dragon.lua's mesh transform keeps
    it inside the scene's walls)."""
    rs = np.random.RandomState(seed)
    u = np.linspace(0.0, 2.0 * np.pi, n_along, endpoint=False)
    v = np.linspace(0.0, 2.0 * np.pi, n_around, endpoint=False)
    p, q = 2.0, 3.0
    r = np.cos(q * u) + 2.2
    c = np.stack([r * np.cos(p * u), r * np.sin(p * u), -np.sin(q * u)], axis=1)  # knot centre line
    t = np.gradient(c, axis=0)
    t = np.concatenate([c[1:] - c[:-1], (c[0] - c[-1])[None]], axis=0) + np.concatenate([(c[0] - c[-1])[None], c[1:] - c[:-1]], axis=0)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    ref = np.array([0.0, 0.0, 1.0])
    n1 = np.cross(t, ref)
    n1 /= np.linalg.norm(n1, axis=1, keepdims=True)
    n2 = np.cross(t, n1)
    # smooth multi-octave displacement of the tube radius (periodic in both parameters)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    rad = np.full_like(uu, 0.55)
    for octave in range(6):
        fu = rs.randint(3, 40) * (octave + 1)
        fv = rs.randint(1, 5) * (octave + 1)
        ph1, ph2 = rs.uniform(0, 2 * np.pi, 2)
        rad += (0.10 / (octave + 1)) * np.sin(fu * uu + ph1) * np.cos(fv * vv + ph2)
    rad += 0.004 * rs.standard_normal(rad.shape)  # fine roughness, so triangles are not coplanar strips
    pos = c[:, None, :] + rad[:, :, None] * (np.cos(vv)[:, :, None] * n1[:, None, :] + np.sin(vv)[:, :, None] * n2[:, None, :])
    # scale into the target box
    lo, hi = pos.reshape(-1, 3).min(0), pos.reshape(-1, 3).max(0)
    target = np.array([8.0, 19.0, 14.0])
    # the knot is widest in x/y and flat in z: stand it up (swap y and z extents by rotating about x)
    pos = pos[:, :, [0, 2, 1]] * np.array([1.0, 1.0, 1.0])
    lo, hi = pos.reshape(-1, 3).min(0), pos.reshape(-1, 3).max(0)
    pos = (pos - lo) / (hi - lo) * target + np.array([-target[0] / 2, -target[1] / 2, 0.0])
    pos = pos.astype(np.float32)
    # smooth normals from the grid tangents (periodic differences), oriented outwards
    du = np.roll(pos, -1, axis=0) - np.roll(pos, 1, axis=0)
    dv = np.roll(pos, -1, axis=1) - np.roll(pos, 1, axis=1)
    nrm = np.cross(du, dv)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=2, keepdims=True), 1e-20)
    centre = pos.mean(axis=1, keepdims=True)
    flip = np.sign(np.sum(nrm * (pos - centre), axis=2, keepdims=True).mean())
    nrm = (nrm * (flip if flip != 0 else 1.0)).astype(np.float32)
    idx = (np.arange(n_along)[:, None] * n_around + np.arange(n_around)[None, :])
    i00 = idx
    i10 = np.roll(idx, -1, axis=0)
    i01 = np.roll(idx, -1, axis=1)
    i11 = np.roll(i10, -1, axis=1)
    tri_v = np.concatenate([np.stack([i00, i10, i11], axis=2).reshape(-1, 3), np.stack([i00, i11, i01], axis=2).reshape(-1, 3)], axis=0)
    tris = np.full((len(tri_v), 3, 3), -1, dtype=np.int32)
    tris[:, :, 0] = tri_v
    tris[:, :, 2] = tri_v
    return Mesh(pos.reshape(-1, 3), np.zeros((0, 2), np.float32), nrm.reshape(-1, 3), [("dragon", tris)])


def grid_mesh(n: int = 32, name: str = "grid", z_amp: float = 0.3, seed: int = 7) -> Mesh:
    """Small bumpy height-field (2*n*n triangles) with UVs and without normals; used by unit tests."""
    rs = np.random.RandomState(seed)
    xs = np.linspace(-1, 1, n + 1)
    xx, yy = np.meshgrid(xs, xs, indexing="ij")
    zz = z_amp * np.sin(3 * xx + rs.uniform(0, 6)) * np.cos(2 * yy + rs.uniform(0, 6)) + 0.02 * rs.standard_normal(xx.shape)
    pos = np.stack([xx, yy, zz], axis=2).reshape(-1, 3).astype(np.float32)
    uv = np.stack([(xx + 1) / 2, (yy + 1) / 2], axis=2).reshape(-1, 2).astype(np.float32)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    tri_v = np.concatenate([np.stack([a, b, c], axis=2).reshape(-1, 3), np.stack([a, c, d], axis=2).reshape(-1, 3)], axis=0)
    tris = np.full((len(tri_v), 3, 3), -1, dtype=np.int32)
    tris[:, :, 0] = tri_v
    tris[:, :, 1] = tri_v
    return Mesh(pos, uv, np.zeros((0, 3), np.float32), [(name, tris)])


# --------------------------------------------------------------------------- reference example scenes
def _cornell_materials():
    s = _spectra()
    colors = {k: spectrum(s[k]) for k in ("white", "green", "red")}
    lamp_color = spectrum(s["lamp"])
    light_m = {"surface": material.emissive(color=lamp_color * 3) + material.diffuse(color=0.78)}
    white = {"surface": material.diffuse(color=colors["white"])}
    green = {"surface": material.diffuse(color=colors["green"])}
    red = {"surface": material.diffuse(color=colors["red"])}
    return dict(light=light_m, left=red, right=green, tall=white, short=white, back=white, ceiling=white, floor=white)


def _julia_object():
    """pyrite/test/cornell/cornell.lua:53-73"""
    return shape.ray_marched(
        shape=ray_marched.quaternion_julia(iterations=25, threshold=4, constant=vector(-0.2, 0.8, 0, 0), slice_plane=0,
                                           variant=quaternion_julia.cubic),
        bounds=bounds.box(min=vector(-7, -1, 0), max=vector(-1, 2, 2)),
        material={"surface": mix(material.mirror(color=1), material.diffuse(color=0.8), fresnel(1.5))},
    )


def cornell(width=512, height=512, spp=64, integrator="simple", fractal=False, spectrum_samples=10, bounces=4, light_samples=1,
            light_bounces=4):
    """pyrite/test/cornell/cornell.lua; defaults are BASELINE config C1 (simple, 512x512x64, S=10, B=4, L=1)."""
    make = renderer.simple if integrator == "simple" else renderer.bidirectional
    objects = [shape.mesh(file=_mesh("box"), materials=_cornell_materials())]
    if fractal:
        objects.append(_julia_object())
    r = make(pixel_samples=spp, spectrum_samples=spectrum_samples, spectrum_bins=50, tile_size=32, light_samples=light_samples,
             bounces=bounces)
    if integrator != "simple":
        r["light_bounces"] = light_bounces
    return {
        "image": {"width": width, "height": height, "white": blackbody(4000)},
        "renderer": r,
        "camera": camera.perspective(fov=37.7, transform=transform.look_at(**{"from": vector(-2.78, -8, 2.73), "to": vector(-2.78, 0, 2.73),
                                                                                "up": vector(z=1)})),
        "world": {"objects": objects},
    }


def dragon(width=1920, height=1080, spp=256, integrator="simple", glass=False, mesh=None, bounces=8, light_samples=4,
           spectrum_samples=10, light_bounces=8):
    """pyrite/test/dragon/dragon.lua with the stand-in mesh; defaults are BASELINE config C2
    (simple, 1920x1080x256, diffuse/glossy material, S=10, B=8, L=4)."""
    make = renderer.simple if integrator == "simple" else renderer.bidirectional
    if glass:
        surface = material.refractive(ior=1.5, _ior=2.37782, dispersion=0.01371, color=1)
    else:
        surface = mix(material.mirror(color=1), material.diffuse(color=0.6), fresnel(1.5))
    r = make(pixel_samples=spp, spectrum_samples=spectrum_samples, spectrum_bins=50, tile_size=32, bounces=bounces,
             light_samples=light_samples)
    if integrator != "simple":
        r["light_bounces"] = light_bounces
    return {
        "image": {"width": width, "height": height},
        "renderer": r,
        "camera": camera.perspective(fov=27, transform=transform.look_at(**{"from": vector(x=-40, y=-30, z=20), "to": vector(z=4.7),
                                                                              "up": vector(z=1)})),
        "world": {"objects": [
            shape.mesh(file=mesh if mesh is not None else dragon_mesh(), materials={"dragon": {"surface": surface}},
                       transform=transform.look_at(**{"from": vector(), "to": vector(0, 0, -1), "up": vector(8, 2, 0)})),
            shape.plane(origin=vector(), normal=vector(z=1), material={"surface": material.diffuse(color=0.4)}),
            shape.plane(origin=vector(y=-10), normal=vector(y=-1), material={"surface": material.diffuse(color=0.4)}),
            shape.plane(origin=vector(x=-11), normal=vector(x=1), material={"surface": material.diffuse(color=0.4)}),
            light.point(position=vector(x=10, y=-25, z=60), direction=vector(x=-10, y=25, z=-57), beam_angle=6,
                        color=light_source.d65 * 5000, width=0.53),
        ]},
    }


def diamonds(width=512, height=300, spp=200, bounces=256, integrator="simple", light_bounces=8):
    """pyrite/test/diamonds/diamonds.lua (BASELINE config C3 at 1920x1080)."""
    diamond = {"surface": material.refractive(ior=2.37782, dispersion=0.01371, color=1)}
    plexi = {"surface": material.mirror(color=mix(0, 0.2, fresnel(1.1)))}
    return {
        "image": {"width": width, "height": height},
        "renderer": (renderer.simple(pixel_samples=spp, spectrum_samples=1, spectrum_bins=50, tile_size=32, bounces=bounces)
                     if integrator == "simple" else
                     renderer.bidirectional(pixel_samples=spp, spectrum_samples=1, spectrum_bins=50, tile_size=32, bounces=bounces, light_bounces=light_bounces)),
        "camera": camera.perspective(fov=12.5, transform=transform.look_at(**{"from": vector(-6.55068, -8.55076, 4.0),
                                                                                "to": vector(0.1, 0, 0.1), "up": vector(z=1)}),
                                     focus_distance=11.08, aperture=0.02),
        "world": {"objects": [shape.mesh(file=_mesh("diamonds"), materials={
            "diamonds": diamond,
            "light_left": {"surface": material.emissive(color=light_source.d65)},
            "light_right": {"surface": material.emissive(color=light_source.d65 * 2)},
            "bottom": plexi,
        })]},
    }


def spheres(width=512, height=256, spp=600, integrator="simple"):
    """pyrite/test/spheres/spheres.lua"""
    ball = shape.sphere(radius=1.5, position=vector(0, 1.4, 10))
    return {
        "image": {"width": width, "height": height},
        "camera": camera.perspective(fov=53, transform=transform.look_at(**{"from": vector(0, 1, 0), "to": vector(0, 1, 1)})),
        "renderer": (renderer.simple if integrator == "simple" else renderer.bidirectional)(
            pixel_samples=spp, spectrum_samples=10, spectrum_bins=50, tile_size=32, light_samples=4),
        "world": {"objects": [
            shape.sphere(radius=50.0, position=vector(0, -50, 10), material={"surface": material.diffuse(color=1)}),
            ball.with_(position=ball.position.with_(y=1.5), material={"surface": material.emissive(color=light_source.d65 * 3)}),
            ball.with_(position=ball.position.with_(x=-3), material={"surface": mix(
                material.mirror(color=1),
                material.diffuse(color=spectrum(format="curve", points=[(400, 0), (450, 0.3), (500, 0), (550, 1), (600, 0)])),
                fresnel(1.5))}),
            ball.with_(position=ball.position.with_(x=3), material={"surface": material.diffuse(
                color=spectrum(format="curve", points=[(580, 0), (600, 1), (610, 1), (650, 0)]))}),
        ]},
    }


def rgb_emission(width=1024, height=256, spp=500):
    """pyrite/test/rgb_emission/rgb_emission.lua (Burns RGB -> spectrum upsampling)"""
    ball = shape.sphere(radius=1, position=vector(0, 2, 0))
    cols = [(1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 1, 1), (0, 0, 1), (1, 0, 1)]
    xs = [-6.25, -3.75, -1.25, 1.25, 3.75, 6.25]
    objs = [shape.plane(origin=vector(z=1), normal=vector(z=1), material={"surface": material.diffuse(color=0.8)})]
    for c, x in zip(cols, xs):
        objs.append(ball.with_(material={"surface": material.emissive(color=rgb(*c))}, position=ball.position.with_(x=x)))
    return {
        "image": {"width": width, "height": height},
        "renderer": renderer.simple(pixel_samples=spp, spectrum_samples=5, spectrum_bins=50, tile_size=32, light_samples=5),
        "camera": camera.perspective(fov=53, transform=transform.look_at(**{"from": vector(0, 0, 15), "to": vector(0, 0, 0)})),
        "world": {"objects": objs},
    }


def textures(width=1024, height=512, spp=400):
    """pyrite/test/textures/textures.lua; the `fabric` textures are absent from the reference, so
    the cube reuses the tactile-paving set."""
    light_ball = shape.sphere(material={"surface": material.emissive(color=light_source.d65 * 20)}, position=vector(0, 0, 0), radius=1)
    floor_material = {
        "surface": mix(material.mirror(color=1), material.diffuse(color=texture(_image("tiles_color.png"))), fresnel(1.5)),
        "normal_map": texture(_image("tiles_normal.png"), "linear") * vector(1, -1, 1),
    }
    return {
        "image": {"width": width, "height": height},
        "renderer": renderer.simple(pixel_samples=spp, spectrum_samples=10, spectrum_bins=50, tile_size=32, bounces=8, light_samples=2),
        "camera": camera.perspective(fov=53, transform=transform.look_at(**{"from": vector(0, 2, 12), "to": vector(0, 2, 0)})),
        "world": {"objects": [
            shape.plane(origin=vector(), normal=vector(y=1), material=floor_material, texture_scale=5),
            light_ball.with_(position=vector(-1, 12, 2), radius=3),
            light_ball.with_(position=vector(15, 3, 4)),
            shape.mesh(file=_mesh("color_checker"), materials={"color_checker": {"surface": material.diffuse(
                color=texture(_image("color_checker.png")))}}),
            shape.sphere(position=vector(-3, 1, 0), radius=1, texture_scale=vector(0.5, 1), material={
                "surface": material.diffuse(color=texture(_image("tactile_paving_color.png"))),
                "normal_map": texture(_image("tactile_paving_normal.png"), "linear") * vector(1, -1, 1)}),
            shape.mesh(file=_mesh("cube"), transform=transform.look_at(**{"from": vector(2, 0.5, 1), "to": vector(-1, 0.5, 2)}), materials={
                "cube": {"surface": material.diffuse(color=texture(_image("tactile_paving_color.png"))),
                         "normal_map": texture(_image("tactile_paving_normal.png"), "linear") * vector(1, -1, 0.1)}}),
        ]},
    }


def snowflake_mesh_scene(width=256, height=256, spp=16):
    """The snowflake mesh (2000 triangles) under a sky, for ray-batch parity (geometry from pyrite/test/snowflake)."""
    return {
        "image": {"width": width, "height": height},
        "renderer": renderer.simple(pixel_samples=spp, spectrum_samples=4, tile_size=32, bounces=6, light_samples=1),
        "camera": camera.perspective(fov=40, transform=transform.look_at(**{"from": vector(0, -6, 3), "to": vector(0, 0, 0.5), "up": vector(z=1)})),
        "world": {"sky": light_source.d65, "objects": [
            shape.mesh(file=_mesh("snowflake"), materials={"snowflake": {"surface": material.refractive(ior=1.31, color=1)}}),
            shape.plane(origin=vector(z=-2), normal=vector(z=1), material={"surface": material.diffuse(color=0.5)}),
            shape.sphere(position=vector(3, -2, 4), radius=0.7, material={"surface": material.emissive(color=light_source.d65 * 10)}),
        ]},
    }


def fractals(width=3840, height=2160, spp=512, bounces=8, light_samples=4, spectrum_samples=10):
    """BASELINE config C4: Mandelbulb + cubic quaternion Julia, floor plane, one emissive sphere."""
    glossy = {"surface": mix(material.mirror(color=1), material.diffuse(color=0.8), fresnel(1.5))}
    return {
        "image": {"width": width, "height": height},
        "renderer": renderer.simple(pixel_samples=spp, spectrum_samples=spectrum_samples, tile_size=32, bounces=bounces,
                                    light_samples=light_samples),
        "camera": camera.perspective(fov=35, transform=transform.look_at(**{"from": vector(0.5, -7.5, 2.6), "to": vector(0.5, 0, 1.1),
                                                                              "up": vector(z=1)})),
        "world": {"sky": light_source.d65 * 0.15, "objects": [
            shape.ray_marched(shape=ray_marched.mandelbulb(iterations=20, threshold=4, power=8),
                              bounds=bounds.box(min=vector(-1.2 - 1.4, -1.2, 0), max=vector(1.2 - 1.4, 1.2, 2.4)), material=glossy),
            shape.ray_marched(shape=ray_marched.quaternion_julia(iterations=25, threshold=4, constant=vector(-0.2, 0.8, 0, 0), slice_plane=0,
                                                                 variant=quaternion_julia.cubic),
                              bounds=bounds.box(min=vector(0.6, -1.5, 0), max=vector(3.6, 1.5, 2.4)), material=glossy),
            shape.plane(origin=vector(), normal=vector(z=1), material={"surface": material.diffuse(color=0.5)}),
            shape.sphere(position=vector(-2, -3, 6), radius=1.2, material={"surface": material.emissive(color=light_source.d65 * 12)}),
        ]},
    }


def fractal_variants(width=512, height=256, spp=64, bounces=4, light_samples=2, spectrum_samples=6):
    """The estimator variants no shipped scene uses: quaternion Julia `regular` and `bicomplex` (distance_estimators.rs:78-107),
    a Mandelbulb with `constant` (the Julia-bulb branch, distance_estimators.rs:16-19), a `bounds.sphere` volume, and an
    `image.filter` expression next to `image.white` (main.rs:197-238)."""
    glossy = {"surface": mix(material.mirror(color=1), material.diffuse(color=0.8), fresnel(1.5))}
    matte = {"surface": material.diffuse(color=spectrum(format="curve", points=[(400, 0.2), (520, 0.9), (700, 0.4)]))}
    return {
        "image": {"width": width, "height": height, "white": blackbody(5000),
                  "filter": spectrum(format="curve", points=[(370, 0.9), (480, 1.0), (600, 0.6), (790, 0.35)])},
        "renderer": renderer.simple(pixel_samples=spp, spectrum_samples=spectrum_samples, tile_size=32, bounces=bounces,
                                    light_samples=light_samples),
        "camera": camera.perspective(fov=38, transform=transform.look_at(**{"from": vector(0.0, -8.5, 2.4), "to": vector(0.0, 0, 1.2),
                                                                              "up": vector(z=1)})),
        "world": {"sky": light_source.d65 * 0.2, "objects": [
            shape.ray_marched(shape=ray_marched.quaternion_julia(iterations=16, threshold=4, constant=vector(-0.291, -0.399, 0.339, 0.437),
                                                                 slice_plane=0.1, variant=quaternion_julia.regular),
                              bounds=bounds.box(min=vector(-4.4, -1.4, 0), max=vector(-1.6, 1.4, 2.8)), material=glossy),
            shape.ray_marched(shape=ray_marched.quaternion_julia(iterations=12, threshold=4, constant=vector(-0.4, 0.45, 0.1, 0.2),
                                                                 slice_plane=0, variant=quaternion_julia.bicomplex),
                              bounds=bounds.sphere(position=vector(0, 0, 1.3), radius=1.3), material=matte),
            shape.ray_marched(shape=ray_marched.mandelbulb(iterations=12, threshold=4, power=6, constant=vector(0.35, -0.45, 0.3)),
                              bounds=bounds.box(min=vector(1.7, -1.2, 0), max=vector(4.1, 1.2, 2.4)), material=glossy),
            shape.plane(origin=vector(), normal=vector(z=1), material={"surface": material.diffuse(color=0.5)}),
            shape.sphere(position=vector(-1, -4, 6), radius=1.0, material={"surface": material.emissive(color=light_source.d65 * 14)}),
        ]},
    }


def bdpt_cornell_dragon(width=3840, height=2160, spp=1024, bounces=6, light_bounces=6, light_samples=1, mesh=None, dragon_scale=0.22):
    """BASELINE config C5: textured Cornell box (tiles floor, colour-checker back wall) + dragon stand-in, bidirectional."""
    mats = _cornell_materials()
    mats["floor"] = {"surface": material.diffuse(color=texture(_image("tiles_color.png")))}
    mats["back"] = {"surface": material.diffuse(color=texture(_image("color_checker.png")))}
    glossy = mix(material.mirror(color=1), material.diffuse(color=0.6), fresnel(1.5))
    return {
        "image": {"width": width, "height": height, "white": blackbody(4000)},
        "renderer": renderer.bidirectional(pixel_samples=spp, spectrum_samples=10, tile_size=32, light_samples=light_samples,
                                           bounces=bounces, light_bounces=light_bounces),
        "camera": camera.perspective(fov=37.7, transform=transform.look_at(**{"from": vector(-2.78, -8, 2.73), "to": vector(-2.78, 0, 2.73),
                                                                                "up": vector(z=1)})),
        "world": {"objects": [
            shape.mesh(file=_mesh("box"), materials=mats),
            shape.mesh(file=mesh if mesh is not None else dragon_mesh(), materials={"dragon": {"surface": glossy}}, scale=dragon_scale,
                       transform=transform.look_at(**{"from": vector(2.9, -1.6, 0.02), "to": vector(2.9, -1.6, -1), "up": vector(8, 2, 0)})),
        ]},
    }


SCENES = {
    "cornell": cornell, "dragon": dragon, "diamonds": diamonds, "spheres": spheres, "rgb_emission": rgb_emission,
    "textures": textures, "snowflake": snowflake_mesh_scene, "fractals": fractals, "bdpt_cornell_dragon": bdpt_cornell_dragon,
    "fractal_variants": fractal_variants,
}
