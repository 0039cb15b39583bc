"""The oracle against closed forms, physical invariants and the committed golden vectors.
The reference ships no tests or golden vectors, so these are what pins the oracle."""
import json
from pathlib import Path

import numpy as np
import pytest
from conftest import scene_ir
from oracle_lib import HIT_DTYPE, RAY_DTYPE, Oracle

from pyrite_b200 import project as P

GOLDEN = Path(__file__).resolve().parent / "golden"


def rays_from(o, d):
    r = np.zeros(len(o), RAY_DTYPE)
    r["o"] = o
    r["d"] = d
    return r


def one_object_scene(obj, **renderer):
    r = dict(pixel_samples=1, light_samples=0, spectrum_samples=2)
    r.update(renderer)
    return P.serialize_project({
        "image": {"width": 8, "height": 8}, "renderer": P.renderer.simple(**r),
        "camera": P.camera.perspective(fov=40, transform=P.transform.look_at(**{"from": P.vector(0, 0, 5), "to": P.vector(0, 0, 0)})),
        "world": {"sky": 1.0, "objects": [obj]}})


def test_sphere_hits_match_the_closed_form():
    ir = one_object_scene(P.shape.sphere(position=P.vector(0.5, -0.25, 1.0), radius=2.0, material={"surface": P.material.diffuse(color=1)}))
    o = Oracle(ir)
    rs = np.random.RandomState(1)
    origin = rs.uniform(-6, 6, (2000, 3)).astype(np.float32)
    origin = origin[np.linalg.norm(origin - [0.5, -0.25, 1.0], axis=1) > 2.5]
    d = rs.standard_normal((len(origin), 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    hits, _ = o.trace(rays_from(origin, d))
    oc = origin.astype(np.float64) - [0.5, -0.25, 1.0]
    b = np.sum(oc * d, axis=1)
    disc = b * b - (np.sum(oc * oc, axis=1) - 4.0)
    t = -b - np.sqrt(np.maximum(disc, 0))
    expect = (disc > 1e-3) & (t > 1e-3)
    assert np.array_equal(hits["kind"][expect], np.full(expect.sum(), 3))
    assert np.allclose(hits["t"][expect], t[expect], rtol=2e-4)
    assert np.all(hits["kind"][disc < -1e-3] == 0)


def test_plane_position_is_mirrored_quirk_q1():
    # collision's Plane stores d = p.n and intersects n.x = -d
    ir = one_object_scene(P.shape.plane(origin=P.vector(0, 0, -2), normal=P.vector(0, 0, 1), material={"surface": P.material.diffuse(color=1)}))
    o = Oracle(ir)
    hits, _ = o.trace(rays_from(np.array([[0, 0, 5]], np.float32), np.array([[0, 0, -1]], np.float32)))
    assert hits["kind"][0] == 1 and hits["t"][0] == pytest.approx(3.0)  # surface at z = +2, not -2


def test_triangle_barycentrics_and_tie_rule():
    # two coincident triangles: the earlier leaf in BVH pre-order wins (world.rs:288-296 strict `<`)
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    tri = np.array([[[0, -1, -1], [1, -1, -1], [2, -1, -1]]], np.int32)
    mesh = P.Mesh(pos, np.zeros((0, 2)), np.zeros((0, 3)), [("a", tri), ("b", tri)])
    m = {"surface": P.material.diffuse(color=1)}
    ir = one_object_scene(P.shape.mesh(file=mesh, materials={"a": m, "b": m}))
    o = Oracle(ir)
    hits, _ = o.trace(rays_from(np.array([[0.25, 0.5, 1]], np.float32), np.array([[0, 0, -1]], np.float32)))
    first_leaf = o.bvh_leaf_order()[0]
    assert hits["kind"][0] == 2 and hits["prim_id"][0] == first_leaf
    assert hits["u"][0] == pytest.approx(0.25) and hits["v"][0] == pytest.approx(0.5) and hits["t"][0] == pytest.approx(1.0)


def test_furnace_white_sphere_in_uniform_sky():
    """A white diffuse sphere inside a uniform sky of radiance 1 must look like the sky (energy conservation
    of the uniform-hemisphere sampling with brdf = 2|n.w|, materials/diffuse.rs)."""
    ir = P.serialize_project({
        "image": {"width": 16, "height": 16}, "renderer": P.renderer.simple(pixel_samples=400, light_samples=0, spectrum_samples=2, bounces=64),
        "camera": P.camera.perspective(fov=20, transform=P.transform.look_at(**{"from": P.vector(0, 0, 5), "to": P.vector(0, 0, 0)})),
        "world": {"sky": 1.0, "objects": [P.shape.sphere(position=P.vector(0, 0, 0), radius=1.0, material={"surface": P.material.diffuse(color=1)}),
                                          P.light.point(position=P.vector(0, 50, 0), color=0)]}})  # a lamp must exist: pick_lamp panics otherwise
    o = Oracle(ir)
    o.render(seed=3)
    film = o.film()
    mean = film[4:12, 4:12, :, 0].sum() / film[4:12, 4:12, :, 1].sum()
    assert mean == pytest.approx(1.0, abs=0.02)


def test_film_expose_and_develop_quirks():
    o = Oracle(scene_ir("cornell"))
    i = o.info
    o.set_film(np.zeros((i.height, i.width, i.bins, 2), np.float32))
    o.expose(np.array([[0.0, 0.0], [-1.0, -1.0], [0.999, 0.999], [0.0, 1.5]], np.float32),
             np.array([[2.0, 380.0, 1.0], [1.0, 779.9, 0.5], [3.0, 500.0, 1.0], [9.0, 500.0, 1.0]], np.float32))
    f = o.film()
    assert f[32, 32, 0, 0] == 2.0 and f[32, 32, 0, 1] == 1.0            # bin 0 at 380 nm
    assert f[0, 0, 63, 0] == 0.5 and f[0, 0, 63, 1] == 0.5              # value * weight, last bin
    assert f[63, 63, 19, 0] == 3.0                                      # (500-380)*64/400 = 19.2
    assert f[..., 1].sum() == 2.5                                       # the sample outside the film is dropped
    film = np.ones((i.height, i.width, i.bins, 2), np.float32)
    o.set_film(film)
    xyz, srgb = o.develop()
    assert np.all(xyz[-1, -1] == 0) and np.all(srgb[-1, -1] == 0)        # film.rs:299 off-by-one: last pixel never developed
    assert np.all(xyz[0, 0] > 0)


def test_golden_vectors():
    g = json.loads((GOLDEN / "oracle_vectors.json").read_text())
    for name, rec in g["scenes"].items():
        o = Oracle(scene_ir(name))
        for kind in ("0", "1", "2"):
            rays = o.gen_rays(int(kind), rec["n_rays"], seed=int(kind))
            hits, _ = o.trace(rays)
            assert hits["prim_id"].tolist() == rec["hits"][kind]["prim_id"], f"{name} batch {kind}: ids drifted"
            assert np.allclose(hits["t"][hits["kind"] != 0], np.array(rec["hits"][kind]["t"], np.float32), rtol=1e-5)
        assert o.bvh_leaf_order()[:32].tolist() == rec["leaf_order_head"]
        pos, ray, wl, hero = o.camera_sample(7, 1, 5)
        assert np.allclose(pos, rec["camera"]["pos"], rtol=1e-6) and hero == rec["camera"]["hero"]
        assert np.allclose(wl, rec["camera"]["wl"], rtol=1e-6)
        o.render(seed=11, threads=1)
        film = o.film()
        assert float(film[..., 1].sum()) == rec["film_weight_sum"]
        assert float(film[..., 0].astype(np.float64).sum()) == pytest.approx(rec["film_acc_sum"], rel=2e-3)


def test_errors_mirror_the_reference():
    from oracle_lib import OracleError

    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    tri = np.array([[[0, -1, -1], [1, -1, -1], [2, -1, -1]]], np.int32)
    mesh = P.Mesh(pos, np.zeros((0, 2)), np.zeros((0, 3)), [("a", tri)])
    with pytest.raises(OracleError, match="missing material"):   # world.rs:194-209
        Oracle(one_object_scene(P.shape.mesh(file=mesh, materials={})))
    with pytest.raises(OracleError, match="vector"):            # a vector where a number is needed
        Oracle(one_object_scene(P.shape.sphere(position=P.vector(), radius=1, material={"surface": P.material.diffuse(color=P.vector(1, 1, 1))})))


def test_hit_dtype_sizes():
    assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 20


def test_eager_emissive_draw_is_distribution_neutral():
    """The wavefront draws a shape lamp's emissive-component index before the visibility result, the reference
    draws it only for unblocked samples (tracer.rs:381-400).  Same distribution, different stream: the oracle in the
    reference's order and in the product's order must agree in mean luminance within the noise of independent renders."""
    from conftest import scene_ir
    from oracle_lib import Oracle

    o = Oracle(scene_ir("spheres", width=32, height=24))   # one emissive sphere is the lamp
    means = {}
    for eager in (True, False):
        ys = []
        for seed in range(12):
            o.render(seed=100 + seed, spp=64, eager_emissive_draw=eager, threads=4)
            y = o.develop()[0][..., 1]
            ys.append(float(np.minimum(y, 5 * y.mean()).mean()))
        means[eager] = (np.mean(ys), np.std(ys) / np.sqrt(len(ys)))
    diff = abs(means[True][0] - means[False][0])
    sigma = float(np.hypot(means[True][1], means[False][1]))
    assert diff <= 4 * sigma and diff <= 5e-3 * means[False][0], (means, diff, sigma)
