"""Product host logic + device stage functions, emulated on the CPU, against the oracle.
The emulation (tests/host_emu.cpp) compiles the product's scene builder and the very same
__host__ __device__ code the kernels run, without FMA, so results must be bit-identical."""
import numpy as np
import pytest
from conftest import BIDIR_NAMES, SCENE_NAMES, scene_ir

from pyrite_b200 import project as P


@pytest.mark.parametrize("name", SCENE_NAMES)
def test_scene_build_matches_oracle(name, emu_factory):
    emu, oracle = emu_factory(name)
    i = oracle.info
    assert emu.info["n_objects"] == i.n_objects and emu.info["n_planes"] == i.n_planes and emu.info["n_lamps"] == i.n_lights
    assert emu.info["n_materials"] == i.n_materials
    # 4-wide nodes: the binary tree (n - 1 interior nodes) folded two levels at a time
    assert (i.n_objects - 1 + 2) // 3 <= emu.info["n_nodes"] <= max(i.n_objects - 1, 0) and i.n_bvh_nodes == max(2 * i.n_objects - 1, 0)
    assert np.array_equal(emu.leaf_order(), oracle.bvh_leaf_order()), "BVH leaf pre-order (the tie rule) differs"


@pytest.mark.parametrize("name", SCENE_NAMES)
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_ray_batches_bit_exact(name, kind, emu_factory):
    emu, oracle = emu_factory(name)
    rays = oracle.gen_rays(kind, 4000, seed=10 + kind)
    want, cw = oracle.trace(rays)
    got, cg = emu.trace(rays)
    for f in ("prim_id", "kind", "t", "u", "v"):
        assert np.array_equal(want[f], got[f]), f"{name} batch {kind}: field {f} differs"
    assert cg["leaves"] <= cw["leaves"] * 1.6 + 100  # nearest-first order must not test many more leaves than the reference order


@pytest.mark.parametrize("name", SCENE_NAMES + BIDIR_NAMES)
def test_film_bit_exact_on_identical_streams(name, emu_factory):
    emu, oracle = emu_factory(name)
    oracle.counters(reset=True)
    oracle.render(seed=5, threads=2)
    emu.render(seed=5)
    fo, fe = oracle.film(), emu.film()
    if name in BIDIR_NAMES:  # fractional weights: the add order inside a bin differs
        assert np.allclose(fo[..., 1], fe[..., 1], rtol=1e-5, atol=1e-6)
        assert oracle.counters(reset=True)["rays"] == emu.rays_last
    else:
        assert np.array_equal(fo[..., 1], fe[..., 1])
    assert np.allclose(fo[..., 0], fe[..., 0], rtol=1e-5, atol=1e-5)  # float add order inside one bin only
    xo, so = oracle.develop()
    xe, se = emu.develop()
    assert np.allclose(xo, xe, rtol=1e-5, atol=1e-7)
    assert np.abs(so.astype(int) - se.astype(int)).max() <= 1


def test_sharded_render_equals_single(emu_factory):
    emu, oracle = emu_factory("cornell")
    emu.render(seed=9)
    whole = emu.film()
    emu.render(seed=9, sample_offset=0, sample_stride=2, reset_film=True)
    emu.render(seed=9, sample_offset=1, sample_stride=2, reset_film=False)
    parts = emu.film()
    assert np.array_equal(whole[..., 1], parts[..., 1])
    assert np.allclose(whole[..., 0], parts[..., 0], rtol=1e-5, atol=1e-6)


def test_camera_samples_match(emu_factory):
    for name in ("cornell", "diamonds"):  # pinhole and thin lens
        emu, oracle = emu_factory(name)
        for tile, sample in [(0, 0), (1, 17), (3, 4095)]:
            po, ro, wo, ho = oracle.camera_sample(3, tile, sample)
            pe, re_, we, he = emu.camera_sample(3, tile, sample, oracle.info.spectrum_samples)
            assert np.array_equal(po, pe) and ho == he and np.array_equal(wo, we)
            assert np.array_equal(ro["o"], re_["o"]) and np.array_equal(ro["d"], re_["d"])


def test_builder_errors():
    from emu_lib import Emu, EmuError

    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    tri = np.array([[[0, -1, -1], [1, -1, -1], [2, -1, -1]]], np.int32)
    mesh = P.Mesh(pos, np.zeros((0, 2)), np.zeros((0, 3)), [("a", tri)])
    base = {"image": {"width": 8, "height": 8}, "renderer": P.renderer.simple(pixel_samples=1),
            "camera": P.camera.perspective(fov=40, transform=P.transform.look_at(**{"from": P.vector(0, 0, 5), "to": P.vector(0, 0, 0)}))}
    with pytest.raises(EmuError, match="missing material"):
        Emu(P.serialize_project(dict(base, world={"objects": [P.shape.mesh(file=mesh, materials={})]})), (8, 8, 64))
    with pytest.raises(EmuError, match="vector"):
        Emu(P.serialize_project(dict(base, world={"objects": [P.shape.sphere(position=P.vector(), radius=1, material={
            "surface": P.material.diffuse(color=P.vector(1, 1, 1))})]})), (8, 8, 64))
    with pytest.raises(EmuError, match="photon"):
        Emu(P.serialize_project(dict(base, renderer=P.renderer.photon_mapping(pixel_samples=1), world={"objects": []})), (8, 8, 64))
    with pytest.raises(EmuError):
        Emu(b"\x00" * 16, (8, 8, 64))
    with pytest.raises(EmuError, match="truncated"):
        Emu(scene_ir("cornell")[:-8], (64, 64, 64))


def test_corrupt_ir_is_rejected_before_allocating():
    """A count field larger than the blob must fail as `truncated` before any table is allocated, and a cyclic mix / add
    surface graph (only a hand-made blob can hold one) must fail instead of looping."""
    import struct

    from emu_lib import Emu, EmuError

    base = {"image": {"width": 8, "height": 8}, "renderer": P.renderer.simple(pixel_samples=1),
            "camera": P.camera.perspective(fov=40, transform=P.transform.look_at(**{"from": P.vector(0, 0, 5), "to": P.vector(0, 0, 0)}))}
    surface = P.mix(P.material.mirror(color=1), P.material.diffuse(color=0.5), 0.25)
    ir = bytearray(P.serialize_project(dict(base, world={"sky": 1.0, "objects": [P.shape.sphere(position=P.vector(), radius=1, material={"surface": surface})]})))
    Emu(bytes(ir), (8, 8, 64))  # sane as written
    huge = bytearray(ir)
    huge[8:12] = struct.pack("<I", 0xFFFFFFFF)  # expression-node count
    with pytest.raises(EmuError, match="truncated"):
        Emu(bytes(huge), (8, 8, 64))
    # walk the surface table (kind, then: emissive/diffuse/mirror = one 12-byte expression; mix = lhs, rhs, expression)
    n_nodes, = struct.unpack_from("<I", ir, 8)
    at = 12
    arity = [4, 3, 2, 3, 3, 2, 1, 0, 0, 0]  # vector rgb binary mix clamp fresnel blackbody spectrum color_texture mono_texture
    for _ in range(n_nodes):
        kind, = struct.unpack_from("<I", ir, at)
        at += 4 + (4 if kind == 2 else 0) + (4 if kind >= 7 else 0) + 12 * arity[kind]
    n_surfaces, = struct.unpack_from("<I", ir, at)
    at, patched = at + 4, False
    for index in range(n_surfaces):
        kind, = struct.unpack_from("<I", ir, at)
        if kind in (0, 1, 2):
            at += 4 + 12
        elif kind == 4:
            struct.pack_into("<I", ir, at + 4, index)  # lhs = the mix node itself
            patched = True
            at += 4 + 8 + 12
        else:
            raise AssertionError(f"unexpected surface kind {kind}")
    assert patched
    with pytest.raises(EmuError, match="cyclic"):
        Emu(bytes(ir), (8, 8, 64))


def test_empty_and_degenerate_scenes():
    from emu_lib import Emu
    from oracle_lib import Oracle

    base = {"image": {"width": 8, "height": 8}, "renderer": P.renderer.simple(pixel_samples=2, light_samples=0),
            "camera": P.camera.perspective(fov=40, transform=P.transform.look_at(**{"from": P.vector(0, 0, 5), "to": P.vector(0, 0, 0)}))}
    ir = P.serialize_project(dict(base, world={"sky": 2.0, "objects": []}))        # no geometry at all: every ray sees the sky
    emu, o = Emu(ir, (8, 8, 64)), Oracle(ir)
    emu.render(seed=1); o.render(seed=1)
    assert np.array_equal(emu.film(), o.film()) and emu.film()[..., 1].sum() > 0
    # coincident triangles (zero-extent centroid hull -> the builder's halving branch, bvh.rs:67-93)
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    tri = np.repeat(np.array([[[0, -1, -1], [1, -1, -1], [2, -1, -1]]], np.int32), 5, axis=0)
    mesh = P.Mesh(pos, np.zeros((0, 2)), np.zeros((0, 3)), [("a", tri)])
    ir = P.serialize_project(dict(base, world={"sky": 1.0, "objects": [P.shape.mesh(file=mesh, materials={"a": {"surface": P.material.mirror(color=1)}})]}))
    emu, o = Emu(ir, (8, 8, 64)), Oracle(ir)
    assert np.array_equal(emu.leaf_order(), o.bvh_leaf_order())
    rays = np.zeros(1, dtype=[("o", np.float32, 3), ("pad0", np.float32), ("d", np.float32, 3), ("pad1", np.float32)])
    rays["o"] = [[0.2, 0.2, 1]]; rays["d"] = [[0, 0, -1]]
    assert emu.trace(rays)[0]["prim_id"][0] == o.trace(rays)[0]["prim_id"][0] == o.bvh_leaf_order()[0]


@pytest.mark.parametrize("name", ["textures", "spheres", "dragon", "lua_orbs"])
def test_path_records_match_the_oracle_bounce_by_bounce(name):
    """tools/first_divergence.py on the CPU emulation of the device code: every bounce of every path sample (hit, incident
    direction, position, normal, scattered direction, RNG state) and the exposed brightness are bit-identical to the
    oracle's.  `textures` holds a cube standing on a plane: hits that tie with the plane within an ulp follow the reference's
    pre-order rule (bvh.rs:213 + world.rs:288-296), which a plain nearest-hit minimum gets wrong about once in 3000 paths."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import first_divergence

    rep = first_divergence.run(name, "emu", "glibc", 3000 if name != "lua_orbs" else 600, 5, 2)
    assert rep["diverged"] == 0, rep


def test_ir_spec_is_sufficient_to_write_a_project_by_hand():
    """include/pyrite_b200_ir.h is the layout a host in another language writes: a blob assembled here with
    nothing but `struct.pack` and that header's rules is byte-identical to what project.serialize_project emits for the
    same project, and the scene builder accepts it."""
    import struct

    from emu_lib import Emu

    from pyrite_b200.project import _Serializer  # only for the constant tables (Burns / XYZ / D65), which are data

    tables = _Serializer(None).tables
    u32 = lambda v: struct.pack("<I", v)
    f32 = lambda v: struct.pack("<f", v)
    num = lambda v: u32(0) + struct.pack("<d", v)
    node = lambda i: u32(1) + u32(i) + u32(0)
    none, some = u32(0), u32(1)
    blob = u32(0x52495950) + u32(1)
    # expression nodes: 0 = vector(0, 0, 5, 0) [camera from], 1 = vector(0, 0, 0, 0) [camera to / sphere position], 2 = vector(3, 4, 5, 0) [lamp position]
    blob += u32(3)
    for xyz in ((0.0, 0.0, 5.0), (0.0, 0.0, 0.0), (3.0, 4.0, 5.0)):
        blob += u32(0) + num(xyz[0]) + num(xyz[1]) + num(xyz[2]) + num(0.0)
    blob += u32(1) + u32(1) + num(0.5)                       # surfaces: 0 = diffuse(color = 0.5)
    blob += u32(0) + u32(0) + u32(0) + u32(0)                # no spectra, colour textures, mono textures, meshes
    for lo, hi, key, width in (("burns_min", "burns_max", "burns_rgb", 3), ("xyz_min", "xyz_max", "xyz", 3), ("illum_min", "illum_max", "d65", 1)):
        data = np.ascontiguousarray(tables[key], np.float32)
        blob += f32(float(tables[lo])) + f32(float(tables[hi])) + u32(data.size // width) + data.tobytes()
    blob += u32(16) + u32(8) + none + none                   # image 16 x 8, no filter, no white
    blob += u32(0) + u32(2)                                  # renderer.simple, pixel_samples = 2
    blob += (u32(0) + u32(0)) + (u32(1) + u32(3)) + (u32(1) + u32(1)) + (u32(0) + u32(0)) * 4   # threads -, bounces 3, light_samples 1, rest default
    blob += node(0) + node(1) + none + num(40.0) + none + none   # camera: look_at(from, to, up = nil), fov 40, no thin lens
    blob += some + num(0.25)                                 # sky = 0.25
    blob += u32(2)
    blob += u32(0) + node(1) + num(1.0) + none + u32(0) + none   # sphere(position, radius 1, no texture_scale, material {surface 0, no normal map})
    blob += u32(5) + node(2) + num(30.0)                     # point light(position, color 30)
    ours = P.serialize_project({
        "image": {"width": 16, "height": 8}, "renderer": P.renderer.simple(pixel_samples=2, bounces=3, light_samples=1),
        "camera": P.camera.perspective(fov=40, transform=P.transform.look_at(**{"from": P.vector(0, 0, 5), "to": P.vector(0, 0, 0)})),
        "world": {"sky": 0.25, "objects": [P.shape.sphere(position=P.vector(0, 0, 0), radius=1, material={"surface": P.material.diffuse(color=0.5)}),
                                           P.light.point(position=P.vector(3, 4, 5), color=30)]}})
    emu = Emu(blob, (8, 16, 64))
    assert emu.info["n_objects"] == 1 and emu.info["n_lamps"] == 1
    emu.render(seed=3)
    assert emu.film()[..., 1].sum() > 0
    # project.py interns the two zero vectors as separate nodes (one per table), so compare through the builder, not bytewise, unless equal
    emu2 = Emu(ours, (8, 16, 64))
    emu2.render(seed=3)
    assert np.array_equal(emu.film(), emu2.film())


def test_parallel_scene_build_equals_the_sequential_one(monkeypatch):
    """Big meshes are converted and their BVH is built by several host threads (scene_build.cpp): subtrees are independent, the
    pieces are stitched in depth-first order, so the leaf pre-order (the tie rule of World::intersect) and every hit must be
    exactly those of the one-thread build - and of the oracle."""
    from emu_lib import Emu
    from oracle_lib import Oracle

    from pyrite_b200 import scenes

    ir = P.serialize_project(scenes.dragon(width=64, height=48, spp=1, mesh=scenes.dragon_mesh(400, 100)))  # 80,000 triangles: above the threshold
    monkeypatch.setenv("PYR_BUILD_THREADS", "4")
    many = Emu(ir, (48, 64, 64))
    monkeypatch.setenv("PYR_BUILD_THREADS", "1")
    one = Emu(ir, (48, 64, 64))
    o = Oracle(ir)
    assert many.info == one.info
    assert np.array_equal(many.leaf_order(), one.leaf_order()) and np.array_equal(many.leaf_order(), o.bvh_leaf_order())
    rays = o.gen_rays(1, 3000, seed=4)
    a, _ = many.trace(rays)
    b, _ = one.trace(rays)
    w, _ = o.trace(rays)
    for f in ("prim_id", "kind", "t", "u", "v"):
        assert np.array_equal(a[f], b[f]) and np.array_equal(a[f], w[f])


def test_a_march_that_cannot_advance_still_ends(emu_factory):
    """shapes/mod.rs:127-135 loops `while total < max { total += DE(p); if DE < EPSILON || total > max { break } }`.  Far from the
    origin a step of >= EPSILON can be smaller than half an ulp of `total`: the sum does not change and the reference spins for
    ever (pyrite itself never returns on such a ray).  Found by a 4K render of config C4: a bounce ray leaving the floor plane
    2100 units out.  Product and oracle end the march there (the only deliberate deviation in the ray queries) and must agree."""
    emu, oracle = emu_factory("fractals")
    rays = np.zeros(2, dtype=[("o", np.float32, 3), ("pad0", np.float32), ("d", np.float32, 3), ("pad1", np.float32)])
    rays["o"] = [[2050.99292, -470.784729, 0.0], [2050.99292, -470.784729, 0.0]]
    rays["d"] = [[-0.974572718, 0.224070147, 0.000765009667], [-0.974572718, 0.224070147, 0.0008]]
    want, _ = oracle.trace(rays, threads=1)
    got, _ = emu.trace(rays)
    for f in ("prim_id", "kind", "t"):
        assert np.array_equal(want[f], got[f]), f
    assert want["kind"][0] == 4 and 2100.0 < want["t"][0] < 2105.0


def test_a_march_to_infinity_is_a_miss(emu_factory):
    """An exactly-zero direction component makes the bounding box' exit distance +inf; a ray that then misses the fractal marches
    off to total = +inf, where Shape::ray_intersect reports a hit (`total <= max`) that World::intersect drops (`distance <
    closest` with closest = +inf, world.rs:273-299).  The product's tie rule must not take inf == inf for a tie."""
    emu, oracle = emu_factory("fractals")
    rays = np.zeros(1, dtype=[("o", np.float32, 3), ("pad0", np.float32), ("d", np.float32, 3), ("pad1", np.float32)])
    rays["o"] = [[-2.3175702, -2.6104007, 0.0]]
    rays["d"] = [[0.0, 0.5699651, 0.8216689]]
    want, _ = oracle.trace(rays, threads=1)
    got, _ = emu.trace(rays)
    assert want["kind"][0] == got["kind"][0] and np.array_equal(want["prim_id"], got["prim_id"])
    assert want["kind"][0] == 0 or np.isfinite(want["t"][0])


def _coinciding_centres_scene():
    """Spheres that share a centre (their boxes' centres coincide: Bvh::new halves such groups by position, bvh.rs:66-88),
    some of them several times over, next to ordinary ones."""
    from pyrite_b200.project import camera, material, renderer, shape, transform, vector

    objs = []
    for k in range(7):
        objs.append(shape.sphere(radius=0.5 + 0.1 * k, position=vector(0, 1, 5), material={"surface": material.diffuse(color=0.5)}))
    for k in range(5):
        objs.append(shape.sphere(radius=0.3, position=vector(-2 + k, 0.3, 4 + 0.5 * k), material={"surface": material.diffuse(color=0.8)}))
    objs += [shape.sphere(radius=0.2, position=vector(2, 2, 6), material={"surface": material.emissive(color=3)})] * 3
    return {"image": {"width": 32, "height": 32},
            "camera": camera.perspective(fov=50, transform=transform.look_at(**{"from": vector(0, 1, 0), "to": vector(0, 1, 1)})),
            "renderer": renderer.simple(pixel_samples=2, spectrum_samples=4, spectrum_bins=10, tile_size=16, light_samples=1),
            "world": {"objects": objs}}


@pytest.mark.parametrize("name", SCENE_NAMES + ["coinciding"])
def test_level_synchronous_bvh_build_is_the_depth_first_tree(name):
    """The GPU builds the BVH level by level (pyrite_b200/csrc/bvh_build_core.hpp, run by the kernels of bvh_build.cu); the same
    functions, run level by level on the CPU, must give the tree of the depth-first builder that mirrors Bvh::new
    (spatial/bvh.rs:13-155): identical 4-wide nodes and identical leaf pre-order."""
    from emu_lib import Emu

    ir = P.serialize_project(_coinciding_centres_scene()) if name == "coinciding" else scene_ir(name)
    a, b = Emu(ir, (8, 8, 1)), Emu(ir, (8, 8, 1), level_sync_bvh=True)
    assert a.info == b.info
    assert np.array_equal(a.leaf_order(), b.leaf_order())
    assert a.bvh_digest() == b.bvh_digest()


def test_level_synchronous_bvh_build_on_a_big_mesh():
    from emu_lib import Emu
    from pyrite_b200 import scenes

    ir = P.serialize_project(scenes.dragon(width=64, height=48, spp=1, mesh=scenes.dragon_mesh(1500, 48)))   # 144,000 triangles
    a, b = Emu(ir, (8, 8, 1)), Emu(ir, (8, 8, 1), level_sync_bvh=True)
    assert a.info["n_objects"] == 144000 and a.info == b.info
    assert a.bvh_digest() == b.bvh_digest()


def _random_sphere_project(seed):
    """Spheres drawn from a coarse grid: many coinciding centres (groups halved by position, bvh.rs:66-88), duplicates, empty buckets,
    groups of two and three."""
    from pyrite_b200.project import camera, material, renderer, shape, transform, vector

    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(2, 400))
    grid = int(rng.integers(1, 6))
    objs = []
    for _ in range(n):
        x, y, z = (float(v) for v in rng.integers(0, grid, 3) * 0.5)
        objs.append(shape.sphere(radius=float(rng.choice([0.1, 0.25, 0.4])), position=vector(x, y, z + 4), material={"surface": material.diffuse(color=0.5)}))
    return n, {"image": {"width": 16, "height": 16},
               "camera": camera.perspective(fov=50, transform=transform.look_at(**{"from": vector(0, 1, 0), "to": vector(0, 1, 1)})),
               "renderer": renderer.simple(pixel_samples=1, spectrum_samples=2, spectrum_bins=4, tile_size=16, light_samples=0),
               "world": {"objects": objs}}


@pytest.mark.parametrize("seed", range(6))
def test_level_synchronous_bvh_build_on_random_item_sets(seed):
    """On random item sets with many ties the level-synchronous build must still be the depth-first tree."""
    from emu_lib import Emu

    n, proj = _random_sphere_project(seed)
    ir = P.serialize_project(proj)
    a, b = Emu(ir, (8, 8, 1)), Emu(ir, (8, 8, 1), level_sync_bvh=True)
    assert a.info == b.info and a.info["n_objects"] == n
    assert np.array_equal(a.leaf_order(), b.leaf_order())
    assert a.bvh_digest() == b.bvh_digest()


@pytest.mark.parametrize("n,same", [(2, False), (2, True), (3, True), (64, True), (65, True)])
def test_level_synchronous_bvh_build_on_tiny_and_degenerate_sets(n, same):
    """Two items (one split), and sets whose centres ALL coincide (every level halves by position, bvh.rs:66-88)."""
    from emu_lib import Emu
    from pyrite_b200.project import camera, material, renderer, shape, transform, vector

    objs = [shape.sphere(radius=0.2 + (0.0 if same else 0.1 * k), position=vector(0 if same else k, 1, 5), material={"surface": material.diffuse(color=0.5)})
            for k in range(n)]
    proj = {"image": {"width": 16, "height": 16},
            "camera": camera.perspective(fov=50, transform=transform.look_at(**{"from": vector(0, 1, 0), "to": vector(0, 1, 1)})),
            "renderer": renderer.simple(pixel_samples=1, spectrum_samples=2, spectrum_bins=4, tile_size=16, light_samples=0),
            "world": {"objects": objs}}
    ir = P.serialize_project(proj)
    a, b = Emu(ir, (8, 8, 1)), Emu(ir, (8, 8, 1), level_sync_bvh=True)
    assert a.info == b.info and a.info["n_objects"] == n
    assert np.array_equal(a.leaf_order(), b.leaf_order())
    assert a.bvh_digest() == b.bvh_digest()
