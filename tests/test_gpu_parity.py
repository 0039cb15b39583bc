"""Parity of the CUDA path (through the C ABI, pyrite_b200.api) with the oracle on a B200.

Bars (BASELINE.json north_star / SURVEY.md §8d):
  * BVH hit / primitive ids bit-exact on identical ray batches, edge-grazing ties counted; hit
    distances within 1e-5 relative (triangles, spheres, planes: in practice bit-exact, the kernels
    are compiled without FMA contraction);
  * films on identical per-path RNG streams: mean luminance (CIE Y) within 1e-3 relative and
    per-pixel RMSE(Y)/mean(Y) <= 1e-2.  The device evaluates the shading-side transcendentals
    (sin / cos / acos / atan2 / exp / pow) in double and rounds once, glibc's float functions differ
    from that in the last bit for a few per cent of the arguments, and a last-bit difference in a
    sampled direction decorrelates the rest of that path.  Every film is therefore compared with BOTH
    oracle variants (oracle/pyro_math.hpp): against the double-rounded one the §8d bars hold as
    stated; against the glibc one (the reference's arithmetic) the RMSE may reach what the two
    oracle variants differ by between themselves, measured in the same test - i.e. the allowance is
    the oracle's own sensitivity to libm's last bit, not a free parameter.  tools/first_divergence.py
    shows the first differing value path by path;
  * sphere-traced shapes (CUDA's float powf / acosf / ... inside the distance estimators, no double
    variant exists): mean <= 1e-2 and RMSE <= 1.5 x the RMSE between two independent-seed oracle renders.
Every observed number is also written to gpurun_out/parity_gpu.json (copied to profiles/parity_r2.json).
"""
import json
import os
from pathlib import Path

import numpy as np
import pytest
from conftest import BIDIR_NAMES, MARCHED_SCENES, SCENE_NAMES, scene_ir

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
RECORD = {}


@pytest.fixture(scope="module", autouse=True)
def parity_record():
    yield RECORD
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    merged = {}
    try:  # a partial run (-k ...) adds to what a full run recorded instead of replacing it
        merged = json.loads((out / "parity_gpu.json").read_text())
    except (OSError, ValueError):
        pass
    for key, value in RECORD.items():
        if isinstance(value, dict) and isinstance(merged.get(key), dict):
            merged[key].update(value)
        else:
            merged[key] = value
    (out / "parity_gpu.json").write_text(json.dumps(merged, indent=1, sort_keys=True, default=float) + "\n")


def film_triplet(name, xg, xo, xd):
    """GPU vs oracle (glibc libm), GPU vs oracle (double-rounded libm), and the two oracle variants against each other."""
    g = dict(zip(("mean_rel", "rmse_rel", "off5pct"), luminance_stats(xo, xg)))
    d = dict(zip(("mean_rel", "rmse_rel", "off5pct"), luminance_stats(xd, xg)))
    f = dict(zip(("mean_rel", "rmse_rel", "off5pct"), luminance_stats(xo, xd)))
    RECORD.setdefault("films", {})[name] = {"gpu_vs_oracle_glibc": g, "gpu_vs_oracle_double": d, "oracle_glibc_vs_oracle_double": f}
    print(f"{name}: GPU vs oracle[double libm] mean {d['mean_rel']:.2e} rmse {d['rmse_rel']:.2e} off {d['off5pct']:.2%} | "
          f"GPU vs oracle[glibc] mean {g['mean_rel']:.2e} rmse {g['rmse_rel']:.2e} off {g['off5pct']:.2%} | "
          f"oracle[glibc] vs oracle[double] rmse {f['rmse_rel']:.2e}")
    return g, d, f


def assert_film_bars(name, g, d, f):
    # Against the oracle that rounds libm like the device does the films agree to float-accumulation level (observed on a B200:
    # RMSE/mean <= 6e-7 on all 15 exact scenes), far inside the SURVEY.md §8d bars (mean 1e-3, RMSE 1e-2)
    assert d["mean_rel"] <= 1e-5 and d["rmse_rel"] <= 1e-4, (name, d)
    # against the reference's glibc arithmetic: no further than the oracle itself moves when only libm's last bit changes
    assert g["mean_rel"] <= max(1e-3, 2.0 * f["mean_rel"]) and g["rmse_rel"] <= max(1e-2, 1.5 * f["rmse_rel"]), (name, g, f)

EXACT_SCENES = ["cornell", "spheres", "diamonds", "textures", "rgb_emission", "snowflake", "dragon", "edge_portrait"]


@pytest.mark.parametrize("name", SCENE_NAMES)
def test_leaf_order_and_info(name, gpu_renderer_factory, oracle_factory):
    r, o = gpu_renderer_factory(name), oracle_factory(name)
    assert np.array_equal(r.bvh_leaf_order(), o.bvh_leaf_order())
    gi, oi = r.info, o.info
    for f in ("width", "height", "bins", "algorithm", "pixel_samples", "bounces", "light_samples", "spectrum_samples", "light_bounces",
              "tile_size", "n_objects", "n_planes", "n_lights", "n_bvh_nodes", "n_materials"):
        assert getattr(gi, f) == getattr(oi, f), f


@pytest.mark.parametrize("name", EXACT_SCENES)
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_ray_batches(name, kind, gpu_renderer_factory, oracle_factory):
    r, o = gpu_renderer_factory(name), oracle_factory(name)
    rays = o.gen_rays(kind, 200_000, seed=20 + kind)
    want, _ = o.trace(rays)
    got = r.trace(rays)
    same = (want["prim_id"] == got["prim_id"]) & (want["kind"] == got["kind"])
    ties = int((~same).sum())
    print(f"{name} batch {kind}: {ties} id mismatches of {len(rays)}")
    assert ties == 0, f"{ties} hit ids differ (edge-grazing ties are expected to be 0 for these batches)"
    hit = want["kind"] != 0
    rel = np.abs(want["t"][hit] - got["t"][hit]) / np.abs(want["t"][hit])
    assert rel.size == 0 or rel.max() <= 1e-5
    assert np.allclose(want["u"], got["u"], atol=1e-6) and np.allclose(want["v"], got["v"], atol=1e-6)


@pytest.mark.parametrize("name", ["fractals", "fractal_variants"])
def test_ray_marched_batches_within_tolerance(name, gpu_renderer_factory, oracle_factory):
    """Sphere tracing calls powf / acosf / atan2f / sinf / cosf / logf per step: tolerance-based parity only.  `fractals` =
    Mandelbulb + cubic Julia (config C4's shapes); `fractal_variants` = Julia `regular` and `bicomplex`, a Mandelbulb with
    `constant`, a spherical bounding volume (distance_estimators.rs:16-19,78-107, shapes/mod.rs:660-682)."""
    r, o = gpu_renderer_factory(name), oracle_factory(name)
    rays = o.gen_rays(0, 100_000, seed=1)
    want, _ = o.trace(rays)
    got = r.trace(rays)
    same = (want["prim_id"] == got["prim_id"]) & (want["kind"] == got["kind"])
    hit = (want["kind"] != 0) & same
    rel = np.abs(want["t"][hit] - got["t"][hit]) / np.abs(want["t"][hit])
    marched = int((want["kind"] == 4).sum())
    print(f"{name}: {marched} ray-marched hits of {len(rays)}; id flips {(~same).mean():.2e}; median rel t err {np.median(rel):.1e}; rel t err > 1e-3: {np.mean(rel > 1e-3):.2e}")
    assert marched > 1000
    assert (~same).mean() <= 2e-3
    assert np.median(rel) <= 1e-6 and np.mean(rel > 1e-3) <= 5e-3


def test_edge_cases(gpu_renderer_factory, oracle_factory):
    r, o = gpu_renderer_factory("cornell"), oracle_factory("cornell")
    from pyrite_b200 import api

    assert len(r.trace(np.zeros(0, api.RAY_DTYPE))) == 0   # empty batch
    rays = o.gen_rays(0, 33, seed=5)            # ragged: not a multiple of the 32-ray packet
    assert np.array_equal(o.trace(rays)[0]["prim_id"], r.trace(rays)["prim_id"])
    rays = o.gen_rays(1, 1, seed=6)
    assert np.array_equal(o.trace(rays)[0]["prim_id"], r.trace(rays)["prim_id"])
    rays = o.gen_rays(0, 64, seed=7)
    rays["d"] = -rays["d"]                       # looking away from the box: all misses
    got = r.trace(rays)
    want = o.trace(rays)[0]
    assert np.array_equal(want["kind"], got["kind"]) and np.array_equal(want["prim_id"], got["prim_id"])
    rays["d"][:, :] = [0, 0, 1]                  # axis-aligned directions: 1/0 = inf in the slab test
    assert np.array_equal(o.trace(rays)[0]["prim_id"], r.trace(rays)["prim_id"])
    rays["d"][:, :] = np.nan                     # NaN directions must not hang or crash
    assert np.array_equal(o.trace(rays)[0]["kind"], r.trace(rays)["kind"])


def luminance_stats(xo, xg):
    yo, yg = xo[..., 1].astype(np.float64), xg[..., 1].astype(np.float64)
    mean = yo.mean()
    return abs(yg.mean() - mean) / mean, np.sqrt(np.mean((yo - yg) ** 2)) / mean, float(np.mean(np.abs(yo - yg) > 0.05 * mean))


@pytest.mark.parametrize("name", SCENE_NAMES)
def test_films_on_identical_streams(name, gpu_renderer_factory, oracle_factory):
    from oracle_lib import Oracle

    spp = 8 if name in MARCHED_SCENES else 16
    r, o = gpu_renderer_factory(name), oracle_factory(name)
    r.render(seed=5, spp=spp)
    o.render(seed=5, spp=spp)
    fg, fo = r.film(), o.film()
    if name not in MARCHED_SCENES:
        assert np.array_equal(fg[..., 1], fo[..., 1]), "per-bin weights (sample counts) differ"
    xg, sg = r.develop()
    xo, so = o.develop()
    dmean, rmse, off = luminance_stats(xo, xg)
    if name not in MARCHED_SCENES:
        od = Oracle(scene_ir(name), "double")
        od.render(seed=5, spp=spp)
        g, d, f = film_triplet(name, xg, xo, od.develop()[0])
        assert_film_bars(name, g, d, f)
        assert np.mean(np.abs(sg.astype(int) - so.astype(int)) > 1) <= 0.01 + 2.0 * f["off5pct"] + (0.02 if f["rmse_rel"] > 1e-3 else 0.0)
        c = r.counters()
        assert c["path_samples"] > 0 and c["rays"] > 0 and c["kernel_launches"] > 0
        return
    print(f"{name}: mean-Y rel err {dmean:.2e}, RMSE/mean {rmse:.2e}, pixels off by >5% {off:.2%}")
    if name in MARCHED_SCENES:
        # the ray-marched normal is a difference of nearly equal distance estimates (shapes/mod.rs:387-405), so device-vs-glibc
        # ULPs decorrelate the bounce directions: compare against the oracle's own noise floor (independent-seed mode)
        o.render(seed=6, spp=spp)
        xo2, _ = o.develop()
        _, floor, _ = luminance_stats(xo, xo2)
        print(f"{name}: oracle-vs-oracle noise floor RMSE/mean {floor:.2e}")
        RECORD.setdefault("films", {})[name] = {"gpu_vs_oracle_glibc": {"mean_rel": dmean, "rmse_rel": rmse, "off5pct": off}, "oracle_independent_seed_rmse_rel": floor}
        assert dmean <= max(1e-2, 2.0 * floor / np.sqrt(xo.shape[0] * xo.shape[1])) and rmse <= 1.5 * floor
    c = r.counters()
    assert c["path_samples"] > 0 and c["rays"] > 0 and c["kernel_launches"] > 0


@pytest.mark.parametrize("name", BIDIR_NAMES)
def test_bidirectional_films_on_identical_streams(name, gpu_renderer_factory, oracle_factory):
    """renderer/bidirectional.rs through the wavefront (lamp subpath, camera subpath, connections, light tracing)."""
    from oracle_lib import Oracle

    r, o = gpu_renderer_factory(name), oracle_factory(name)
    spp = 8
    r.counters(reset=True)
    o.counters(reset=True)
    r.render(seed=5, spp=spp)
    o.render(seed=5, spp=spp)
    xg, sg = r.develop()
    xo, so = o.develop()
    dmean, rmse, off = luminance_stats(xo, xg)
    rays_g, rays_o = r.counters()["rays"], o.counters()["rays"]
    print(f"{name}: rays gpu {rays_g} oracle {rays_o}")
    assert abs(rays_g - rays_o) <= 2e-3 * rays_o
    if name == "bd_cornell_fractal":
        o.render(seed=6, spp=spp)
        _, floor, _ = luminance_stats(xo, o.develop()[0])
        RECORD.setdefault("films", {})[name] = {"gpu_vs_oracle_glibc": {"mean_rel": dmean, "rmse_rel": rmse, "off5pct": off}, "oracle_independent_seed_rmse_rel": floor}
        assert dmean <= 1e-2 and rmse <= 1.5 * floor
        return
    od = Oracle(scene_ir(name), "double")
    od.render(seed=5, spp=spp)
    g, d, f = film_triplet(name, xg, xo, od.develop()[0])
    assert_film_bars(name, g, d, f)


def test_sample_pass_sharding_is_additive(gpu_renderer_factory):
    r = gpu_renderer_factory("cornell")
    r.render(seed=9, spp=4)
    whole = r.film()
    r.render(seed=9, spp=4, sample_offset=0, sample_stride=2, reset_film=True)
    r.render(seed=9, spp=4, sample_offset=1, sample_stride=2, reset_film=False)
    parts = r.film()
    assert np.array_equal(whole[..., 1], parts[..., 1])
    assert np.allclose(whole[..., 0], parts[..., 0], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell", "diamonds", "bd_cornell"])
def test_pool_size_does_not_change_the_job(name, gpu_renderer_factory):
    """Paths in flight only change how the job is cut into wavefront iterations: streams are keyed by sample index, so
    ray counts and film weights are identical and the accumulators differ by float atomics order only."""
    if name not in SCENE_NAMES and name not in BIDIR_NAMES:
        pytest.skip("scene not in this build")
    r = gpu_renderer_factory(name)
    films, rays = [], []
    for pool in (0, 4096, 1000):  # library default, a small pool, a ragged one (rounded up to 1024)
        r.counters(reset=True)
        r.render(seed=21, spp=3, pool_paths=pool)
        c = r.counters()
        films.append(r.film())
        rays.append((c["rays"], c["path_samples"], c["path_rays"]))
    assert rays[0] == rays[1] == rays[2]
    for f in films[1:]:
        if name.startswith("bd_"):  # connection weights 1/(len_camera * len_lamp) are not dyadic: their sums depend on the order too
            assert np.allclose(films[0][..., 1], f[..., 1], rtol=2e-4, atol=1e-5)
        else:
            assert np.array_equal(films[0][..., 1], f[..., 1])
        assert np.allclose(films[0][..., 0], f[..., 0], rtol=2e-4, atol=1e-5)


@pytest.mark.gpu
def test_trace_device_wants_aligned_rays(gpu_renderer_factory):
    import torch
    from pyrite_b200 import api
    r = gpu_renderer_factory("cornell")
    rays = np.zeros(66, api.RAY_DTYPE)
    rays["o"] = (0.0, 0.0, 5.0)
    rays["d"] = (0.0, 0.0, -1.0)
    buf = torch.from_numpy(rays.view(np.uint8).copy()).to("cuda:0")
    hits = torch.zeros(20 * 66, dtype=torch.uint8, device="cuda:0")
    assert buf.data_ptr() % 32 == 0
    r.trace_device(buf.data_ptr(), 64, hits.data_ptr())
    with pytest.raises(api.PyriteError, match="32-byte aligned"):
        r.trace_device(buf.data_ptr() + 16, 64, hits.data_ptr())


def test_film_expose_and_develop(gpu_renderer_factory, oracle_factory):
    r, o = gpu_renderer_factory("cornell"), oracle_factory("cornell")
    rs = np.random.RandomState(4)
    n = 100_000
    pos = rs.uniform(-1.2, 1.2, (n, 2)).astype(np.float32)
    smp = np.stack([rs.uniform(0, 5, n), rs.uniform(380, 779.99, n), rs.uniform(0.1, 1, n)], axis=1).astype(np.float32)
    r.clear_film()
    o.set_film(np.zeros(r.film_shape, np.float32))
    r.expose(pos, smp)
    o.expose(pos, smp)
    fg, fo = r.film(), o.film()
    assert np.array_equal(fg[..., 1] > 0, fo[..., 1] > 0)
    assert np.allclose(fg, fo, rtol=1e-5, atol=1e-6)
    film = rs.uniform(0, 2, r.film_shape).astype(np.float32)
    r.set_film(film)
    o.set_film(film)
    for step in (2.0, 30.0):
        xg, sg = r.develop(step)
        xo, so = o.develop(step)
        assert np.allclose(xg, xo, rtol=2e-5, atol=1e-6)
        assert np.abs(sg.astype(int) - so.astype(int)).max() <= 1
        assert np.all(xg[-1, -1] == 0)  # film.rs:299 quirk


def test_camera_seam(gpu_renderer_factory, oracle_factory):
    for name in ("cornell", "diamonds"):
        r, o = gpu_renderer_factory(name), oracle_factory(name)
        for tile, sample in [(0, 0), (1, 17), (3, 4095)]:
            pg, rg, wg, hg = r.camera_sample(3, tile, sample)
            po, ro, wo, ho = o.camera_sample(3, tile, sample)
            assert np.array_equal(pg, po) and hg == ho and np.array_equal(wg, wo)
            assert np.allclose(rg["o"], ro["o"], rtol=1e-6, atol=1e-7) and np.allclose(rg["d"], ro["d"], rtol=1e-6, atol=1e-7)


def test_errors(gpu_renderer_factory):
    from pyrite_b200 import api
    from pyrite_b200 import project as P

    r = api.Renderer(0)
    with pytest.raises(api.PyriteError) as e:
        r.render()
    assert e.value.status == 3  # PYR_ERR_STATE: no project loaded
    with pytest.raises(api.PyriteError) as e:
        r.load(b"garbage!" * 4)
    assert e.value.status == 1
    ir = P.serialize_project({"image": {"width": 8, "height": 8}, "renderer": P.renderer.simple(pixel_samples=1),
                              "camera": P.camera.perspective(fov=40, transform=P.transform.look_at(**{"from": P.vector(0, 0, 5), "to": P.vector(0, 0, 0)})),
                              "world": {"objects": [P.shape.sphere(position=P.vector(), radius=1, material={"surface": P.material.diffuse(color=1)})]}})
    r.load(ir)
    with pytest.raises(api.PyriteError, match="no lamps"):
        r.render()
    seen = []
    r.load(scene_ir("cornell"))
    with pytest.raises(api.PyriteError, match="cancel"):
        r.render(spp=64, pool_paths=1024, progress=lambda pct, msg: seen.append(pct) or True)
    assert seen
    r.close()


def test_full_size_properties(gpu_renderer_factory):
    """BASELINE-size geometry (871,200 triangles): size-independent properties instead of the oracle."""
    from pyrite_b200 import api, project, scenes

    r = api.Renderer(0)
    r.load(project.serialize_project(scenes.dragon(width=480, height=270, spp=1)))
    order = r.bvh_leaf_order()
    assert np.array_equal(np.sort(order), np.arange(r.info.n_objects))      # a permutation: every triangle is a leaf exactly once
    rs = np.random.RandomState(2)
    n = 1 << 20
    o = np.tile(np.array([[-40, -30, 20]], np.float32), (n, 1))
    tgt = np.stack([rs.uniform(-4, 4, n), rs.uniform(-9, 9, n), rs.uniform(0, 14, n)], axis=1)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.zeros(n, api.RAY_DTYPE)
    rays["o"], rays["d"] = o, d
    a = r.trace(rays)
    b = r.trace(rays)
    assert np.array_equal(a, b)                                               # idempotent / deterministic
    hit = a["kind"] == 2
    assert hit.mean() > 0.2
    # re-tracing from just in front of each hit point must hit the same primitive at the remaining distance
    p = o[hit] + d[hit] * (a["t"][hit] * 0.5)[:, None]
    rays2 = np.zeros(int(hit.sum()), api.RAY_DTYPE)
    rays2["o"], rays2["d"] = p, d[hit]
    c = r.trace(rays2)
    agree = c["prim_id"] == a["prim_id"][hit]
    assert agree.mean() > 0.999
    assert np.allclose(c["t"][agree], (a["t"][hit] * 0.5)[agree], rtol=1e-3)
    # stats pass returns the same hits and non-zero counters
    s = r.trace(rays[:4096], stats=True)
    assert np.array_equal(s, a[:4096])
    cn = r.counters()
    assert cn["nodes_visited"] > 0 and cn["leaves_tested"] > 0
    r.close()


def _nccl_worker(rank, world, port, out_dir):
    import os
    import sys
    from pathlib import Path

    import torch
    import torch.distributed as dist

    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    sys.path.insert(0, str(root / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from conftest import scene_ir

    from pyrite_b200 import api
    from pyrite_b200.distributed import render_sharded

    r = api.Renderer(rank)
    r.load(scene_ir("cornell"))
    xyz, srgb = render_sharded(r, seed=33, spp=8)
    if rank == 0:
        np.save(Path(out_dir) / "film.npy", r.film())
        np.save(Path(out_dir) / "xyz.npy", xyz)
    dist.barrier()
    r.close()
    dist.destroy_process_group()


def test_two_gpu_sharded_render_equals_single(tmp_path, gpu_renderer_factory):
    """Sample-pass sharding over 2 GPUs + one NCCL film reduce == the single-GPU render of the same seed."""
    import socket

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = gpu_renderer_factory("cornell")
    r.render(seed=33, spp=8)
    whole, reduced = r.film(), np.load(tmp_path / "film.npy")
    assert np.array_equal(whole[..., 1], reduced[..., 1])
    assert np.allclose(whole[..., 0], reduced[..., 0], rtol=1e-4, atol=1e-5)
    assert np.allclose(r.develop()[0], np.load(tmp_path / "xyz.npy"), rtol=1e-4, atol=1e-6)


def test_c1_full_config_film(gpu_renderer_factory):
    """BASELINE config C1: pyrite/test/cornell (box.obj, camera-to-light), 512x512 at 64 spp, S=10 - the whole
    configuration on identical per-path streams, GPU vs the oracle on the host cores."""
    from oracle_lib import Oracle

    ir = scene_ir("cornell", width=512, height=512, spp=64)
    r = gpu_renderer_factory("cornell", width=512, height=512, spp=64)
    o = Oracle(ir)
    secs_gpu = r.render(seed=64)
    secs_cpu = o.render(seed=64)
    fg, fo = r.film(), o.film()
    assert np.array_equal(fg[..., 1], fo[..., 1])
    xg, sg = r.develop()
    xo, so = o.develop()
    dmean, rmse, off = luminance_stats(xo, xg)
    print(f"C1 512x512x64spp: GPU {secs_gpu:.3f} s, oracle ({o.threads} threads) {secs_cpu:.1f} s; mean-Y rel err {dmean:.2e}, RMSE/mean {rmse:.2e}, "
          f"sRGB bytes differing by more than 1: {np.mean(np.abs(sg.astype(int) - so.astype(int)) > 1):.2e}")
    assert dmean <= 1e-3 and rmse <= 1e-2
    assert np.mean(np.abs(sg.astype(int) - so.astype(int)) > 1) <= 1e-3
    assert r.counters()["path_samples"] >= 512 * 512 * 64


def test_independent_seed_noise_floor(gpu_renderer_factory, oracle_factory):
    """Independent-seed mode: GPU with one seed vs the oracle with another must differ by no more than
    1.5 x what two oracle runs differ by, and agree in mean luminance within 5e-3 (scaled by the noise of the mean)."""
    r, o = gpu_renderer_factory("cornell"), oracle_factory("cornell")
    spp = 64
    r.render(seed=101, spp=spp)
    xg, _ = r.develop()
    o.render(seed=202, spp=spp)
    xa, _ = o.develop()
    o.render(seed=303, spp=spp)
    xb, _ = o.develop()
    # the estimator is heavy-tailed (rare bright paths), so the per-pixel comparison uses luminance clamped at 5 x the mean
    cap = 5.0 * float(xa[..., 1].mean())
    ya, yb, yg = (np.minimum(x[..., 1], cap) for x in (xa, xb, xg))
    floor = float(np.sqrt(np.mean((ya - yb) ** 2)) / ya.mean())
    rmse = float(np.sqrt(np.mean((ya - yg) ** 2)) / ya.mean())
    dmean = abs(float(yg.mean()) - float(ya.mean())) / float(ya.mean())
    print(f"independent seeds: clamped RMSE/mean gpu-vs-oracle {rmse:.3e}, oracle-vs-oracle {floor:.3e}, mean-Y rel err {dmean:.2e}")
    assert rmse <= 1.5 * floor
    assert dmean <= max(5e-3, 3.0 * floor / np.sqrt(ya.size))


def test_full_size_ray_batches_against_oracle():
    """SURVEY.md §8d ray-batch parity at BASELINE size: the 871,200-triangle stand-in, 3 batches (camera, bounce, shadow
    rays) of 2^22 rays each, ids bit-exact and distances within 1e-5 relative."""
    from oracle_lib import Oracle

    from pyrite_b200 import api, project, scenes

    ir = project.serialize_project(scenes.dragon(width=480, height=270, spp=1))
    o = Oracle(ir)
    r = api.Renderer(0)
    r.load(ir)
    assert np.array_equal(r.bvh_leaf_order(), o.bvh_leaf_order())
    total_ties = 0
    for kind in (0, 1, 2):
        rays = o.gen_rays(kind, 1 << 22, seed=40 + kind)
        want, cw = o.trace(rays)
        got = r.trace(rays)
        same = (want["prim_id"] == got["prim_id"]) & (want["kind"] == got["kind"])
        ties = int((~same).sum())
        total_ties += ties
        hit = (want["kind"] != 0) & same
        rel = np.abs(want["t"][hit] - got["t"][hit]) / np.abs(want["t"][hit])
        print(f"dragon 871,200 tris, batch {kind}: {len(rays)} rays, {ties} id mismatches, max rel t err {rel.max() if rel.size else 0:.1e}, "
              f"reference-order nodes/ray {cw['nodes'] / len(rays):.1f}, leaves/ray {cw['leaves'] / len(rays):.1f}")
        assert rel.size == 0 or rel.max() <= 1e-5
        # an id mismatch is only acceptable as an edge-grazing tie: both sides report the same distance
        bad = ~same
        if bad.any():
            assert np.all(want["kind"][bad] == got["kind"][bad])
            assert np.allclose(want["t"][bad], got["t"][bad], rtol=1e-5, atol=0)
    print(f"edge-grazing ties over {3 << 22} rays: {total_ties}")
    assert total_ties <= 3
    r.close()


def test_cli_renders_a_project_lua(tmp_path):
    """`python -m pyrite_b200 project.lua`: load_project -> parse_project -> render -> develop -> render.png (main.rs:52-332)."""
    import subprocess
    import sys
    from pathlib import Path

    from PIL import Image

    root = Path(__file__).resolve().parent.parent
    out = tmp_path / "render.png"
    res = subprocess.run([sys.executable, "-m", "pyrite_b200", str(root / "tests" / "golden" / "scenes" / "orbs.lua"), "--seed", "1", "--out", str(out),
                          "--spp", "32", "--no-preview"], cwd=root, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "Project loading" in res.stdout and "Rendering" in res.stdout and "Total" in res.stdout
    im = np.asarray(Image.open(out))
    assert im.shape == (48, 96, 3) and im.mean() > 1.0
    # the preview path (main.rs:261-299): a develop at 30 nm steps between wavefront batches, written to the same file
    res = subprocess.run([sys.executable, "-m", "pyrite_b200", str(root / "tests" / "golden" / "scenes" / "orbs.lua"), "--seed", "1", "--out", str(out),
                          "--preview-interval", "0", "--pool", "1024", "--spp", "32"], cwd=root, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    previews = int(res.stdout.rsplit("(", 1)[1].split()[0])
    assert previews >= 2, res.stdout
    im2 = np.asarray(Image.open(out))
    assert im2.shape == im.shape and abs(float(im2.mean()) - float(im.mean())) < 0.15 * float(im.mean())
    bad = subprocess.run([sys.executable, "-m", "pyrite_b200", str(tmp_path / "missing.lua")], cwd=root, capture_output=True, text=True, timeout=120)
    assert bad.returncode == 1 and "error while loading project file" in bad.stderr


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs C2-C5 at their real parameters (resolution, S, B, L, mesh size), a 4-spp fraction of the job on
# identical per-path streams, GPU vs the oracle on the host cores.  (C1 at its full 64 spp: test_c1_full_config_film.)
def _full_config_project(config):
    from pyrite_b200 import scenes

    if config == "C2":   # Stanford-dragon stand-in, 871,200 triangles, 1920x1080, simple, S=10 B=8 L=4
        return scenes.dragon()
    if config == "C3":   # diamonds.lua as shipped (dispersion, S=1, B=256, thin lens) at 1920x1080
        return scenes.diamonds(width=1920, height=1080)
    if config == "C4":   # Mandelbulb + cubic quaternion Julia, 3840x2160
        return scenes.fractals()
    if config == "C5":   # textured Cornell box + dragon stand-in, bidirectional, 3840x2160
        return scenes.bdpt_cornell_dragon()
    raise KeyError(config)


@pytest.mark.parametrize("config", ["C2", "C3", "C4", "C5"])
def test_full_config_films(config):
    from oracle_lib import Oracle

    from pyrite_b200 import api, project, scenes

    spp = 4
    ir = project.serialize_project(_full_config_project(config))
    o = Oracle(ir)
    with api.Renderer(0) as r:
        r.load(ir)
        info = r.info
        secs_gpu = r.render(seed=77, spp=spp)
        cg = r.counters()
        xg, sg = r.develop()
    o.counters(reset=True)
    secs_cpu = o.render(seed=77, spp=spp)
    co = o.counters()
    xo, so = o.develop()
    entry = {"resolution": [int(info.width), int(info.height)], "spp": spp, "objects": int(info.n_objects), "gpu_seconds": secs_gpu, "oracle_seconds": secs_cpu,
             "oracle_threads": o.threads, "rays_gpu": cg["rays"], "rays_oracle": co["rays"]}
    assert cg["path_samples"] == co["path_samples"] == info.width * info.height * spp
    if config == "C4":
        # sphere tracing: tolerance-level parity, measured against the oracle's own independent-seed noise at the same spp
        # (per-pixel noise does not depend on the resolution, so the floor comes from a quarter-resolution pair)
        dmean, rmse, off = luminance_stats(xo, xg)
        small = Oracle(project.serialize_project(scenes.fractals(width=960, height=540)))
        small.render(seed=77, spp=spp)
        xa = small.develop()[0]
        small.render(seed=78, spp=spp)
        _, floor, _ = luminance_stats(xa, small.develop()[0])
        entry.update({"gpu_vs_oracle_glibc": {"mean_rel": dmean, "rmse_rel": rmse, "off5pct": off}, "oracle_independent_seed_rmse_rel": floor})
        RECORD.setdefault("full_configs", {})[config] = entry
        print(f"{config}: {entry}")
        assert abs(cg["rays"] - co["rays"]) <= 5e-3 * co["rays"]
        assert dmean <= 1e-2 and rmse <= 1.5 * floor
        return
    od = Oracle(ir, "double")
    od.render(seed=77, spp=spp)
    g, d, f = film_triplet(config, xg, xo, od.develop()[0])
    entry.update({"gpu_vs_oracle_glibc": g, "gpu_vs_oracle_double": d, "oracle_glibc_vs_oracle_double": f})
    RECORD.setdefault("full_configs", {})[config] = entry
    print(f"{config}: {entry}")
    assert abs(cg["rays"] - co["rays"]) <= 2e-3 * co["rays"]
    assert_film_bars(config, g, d, f)


@pytest.mark.parametrize("name", ["spheres", "diamonds", "colors", "cornell", "snowflake"])
def test_reference_example_images(name):
    """SURVEY.md §8(c) pin (3): the reference's own unmodified project files (fixtures made by tools/reference_images.py), rendered at
    their own spp, against the example images the reference ships - framing and structure only (see the tool's docstring for why
    radiometry cannot be pinned by those images)."""
    import sys

    sys.path.insert(0, str(ROOT / "tools"))
    import reference_images as ri

    stats = ri.compare(name, ri.render(name, "gpu", 0))
    RECORD.setdefault("reference_example_images", {})[name] = stats
    print(f"{name}: {stats}")
    bars = {"spheres": (0.98, None), "diamonds": (0.90, 0.5), "colors": (0.95, None), "cornell": (0.75, 0.6), "snowflake": (None, None)}[name]
    if bars[0] is not None:
        assert stats["pearson_log_luminance"] >= bars[0]
        assert stats["pearson_log_luminance"] > stats["pearson_if_mirrored"]   # a mirrored camera convention fits worse
    if bars[1] is not None:
        assert stats["pearson_if_mirrored"] <= bars[1]


@pytest.mark.parametrize("name", ["textures", "spheres", "dragon", "cornell"])
def test_first_divergence_on_the_gpu(name):
    """tools/first_divergence.py, GPU vs both oracle variants: against the double-rounded libm nothing may differ before the
    exposed brightness (float-level), against glibc's libm the first difference must be an ulp-level value, never a hit id."""
    import sys

    sys.path.insert(0, str(ROOT / "tools"))
    import first_divergence

    n = 4000
    rep_d = first_divergence.run(name, "gpu", "double", n, 5, 2)
    rep_g = first_divergence.run(name, "gpu", "glibc", n, 5, 2)
    RECORD.setdefault("first_divergence", {})[name] = {"vs_oracle_double": rep_d, "vs_oracle_glibc": rep_g}
    print(f"{name}: vs double-rounded libm {rep_d['diverged']} of {n} paths differ {rep_d['first_differing_field']}; "
          f"vs glibc {rep_g['diverged']} of {n} {rep_g['first_differing_field']} max ulps {rep_g['max_ulps_at_first_difference']}")
    structural = [k for k in rep_d["first_differing_field"] if k.split("@")[0] in ("kind", "prim_id", "rng_w", "path_length", "exposed_count")]
    assert not structural, rep_d
    assert rep_d["diverged_fraction"] <= 0.02, rep_d
    first_is_id = sum(v for k, v in rep_g["first_differing_field"].items() if k.split("@")[0] in ("kind", "prim_id"))
    assert first_is_id <= 0.002 * n, rep_g


def test_a_march_that_cannot_advance_still_ends(gpu_renderer_factory, oracle_factory):
    """The reference's sphere-tracing loop never returns when a step of >= EPSILON is absorbed by the rounding of `total`
    (tests/test_host_logic.py has the story); the kernels end the march there and agree with the oracle."""
    r, o = gpu_renderer_factory("fractals"), oracle_factory("fractals")
    from pyrite_b200 import api

    rays = np.zeros(64, api.RAY_DTYPE)
    rays["o"] = (2050.99292, -470.784729, 0.0)
    rays["d"] = (-0.974572718, 0.224070147, 0.000765009667)
    rays["d"][1::2] = (-0.974572718, 0.224070147, 0.0008)
    want, _ = o.trace(rays, threads=1)
    got = r.trace(rays)
    assert np.array_equal(want["kind"], got["kind"]) and np.array_equal(want["prim_id"], got["prim_id"])
    assert np.allclose(want["t"], got["t"], rtol=1e-5)


def test_a_march_to_infinity_is_a_miss(gpu_renderer_factory, oracle_factory):
    """A direction component of exactly 0 makes the box exit distance +inf (BoundingVolume::intersect, shapes/mod.rs:591-658); a
    ray that misses the fractal then marches off to total = +inf, `total <= max` holds and Shape::ray_intersect returns a hit at
    +inf, which World::intersect drops (`distance < closest`, closest starts at +inf).  The wavefront (k_march merges by atomicMin)
    must drop it too - found as NaN pixels in the 512-spp C4 render by tools/find_nonfinite.py + tools/replay_sample.py."""
    from pyrite_b200 import api

    r, o = gpu_renderer_factory("fractals"), oracle_factory("fractals")
    rays = np.zeros(32, api.RAY_DTYPE)
    rays["o"] = (-2.3175702, -2.6104007, 0.0)
    rays["d"] = (0.0, 0.5699651, 0.8216689)
    want, _ = o.trace(rays, threads=1)
    got = r.trace(rays)
    # (the ray grazes the Mandelbulb: whether it hits is libm-level; what must never come back is a hit at +inf)
    assert np.isfinite(got["t"][got["kind"] != 0]).all() and np.isfinite(want["t"][want["kind"] != 0]).all()
    print("march to infinity: oracle kind", want["kind"][0], "gpu kind", got["kind"][0])
    # and through the wavefront: a render leaves no more non-finite bins behind than the oracle's does
    r.render(seed=4242, spp=8)
    o.render(seed=4242, spp=8)
    assert (~np.isfinite(r.film())).sum() <= (~np.isfinite(o.film())).sum()


def _load_both_ways(ir):
    """The same project loaded with the host BVH builder and with the GPU one: (digest, leaf order) of each."""
    from pyrite_b200 import api

    got = {}
    old = os.environ.get("PYR_BVH_BUILD")
    try:
        for how in ("host", "gpu"):
            os.environ["PYR_BVH_BUILD"] = how
            with api.Renderer(0) as r:
                r.load(ir)
                r.load(ir)   # the second load is the one timed: the first pays for the library's lazily loaded kernels
                got[how] = (r.bvh_digest(), r.bvh_leaf_order())
    finally:
        if old is None:
            os.environ.pop("PYR_BVH_BUILD", None)
        else:
            os.environ["PYR_BVH_BUILD"] = old
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["spheres", "cornell", "diamonds", "snowflake", "dragon", "coinciding", "random0", "random1", "random2", "random3"])
def test_gpu_bvh_build_is_the_host_tree(name):
    """bvh_build.cu builds Bvh::new's tree (spatial/bvh.rs:13-155) level by level on the GPU: identical 4-wide nodes and leaf
    pre-order to the depth-first host builder (which tests/test_host_logic.py holds against the oracle's)."""
    from test_host_logic import _coinciding_centres_scene, _random_sphere_project

    from pyrite_b200 import project as P

    if name == "coinciding":
        ir = P.serialize_project(_coinciding_centres_scene())
    elif name.startswith("random"):
        ir = P.serialize_project(_random_sphere_project(int(name[6:]))[1])
    else:
        ir = scene_ir(name)
    got = _load_both_ways(ir)
    (dh, oh), (dg, og) = got["host"], got["gpu"]
    assert not dh["built_on_gpu"] and dg["built_on_gpu"]
    assert np.array_equal(oh, og)
    assert dh["nodes"] == dg["nodes"] and dh["leaf_order"] == dg["leaf_order"]


@pytest.mark.gpu
def test_gpu_bvh_build_at_full_size():
    """Config C5's scene (871,236 items, depth 28): same tree from both builders, and the default load builds it on the GPU."""
    from pyrite_b200 import api, project, scenes

    ir = project.serialize_project(scenes.bdpt_cornell_dragon())
    got = _load_both_ways(ir)
    (dh, oh), (dg, og) = got["host"], got["gpu"]
    print(f"C5 scene load: host BVH build {dh['load_seconds']:.3f} s, GPU BVH build {dg['load_seconds']:.3f} s")
    RECORD["bvh_build"] = ({"items": int(len(oh)), "load_seconds_host_bvh": dh["load_seconds"], "load_seconds_gpu_bvh": dg["load_seconds"]})
    assert np.array_equal(oh, og)
    assert dh["nodes"] == dg["nodes"] and dh["leaf_order"] == dg["leaf_order"]
    with api.Renderer(0) as r:
        r.load(ir)
        assert r.bvh_digest()["built_on_gpu"]
