#!/usr/bin/env python3
"""First-contact GPU probe: smoke check, parity on every example scene, and rough timings on the
C2 dragon scene.  Writes gpurun_out/probe.json."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import __graft_entry__ as g
from oracle_lib import Oracle
from pyrite_b200 import api, project, scenes

out = {}
g.smoke()
small = {"cornell": dict(width=64, height=64, spp=4), "spheres": dict(width=64, height=48, spp=4), "diamonds": dict(width=64, height=48, spp=4),
         "textures": dict(width=64, height=48, spp=4), "rgb_emission": dict(width=64, height=48, spp=4), "snowflake": dict(width=64, height=48, spp=4),
         "fractals": dict(width=48, height=32, spp=2), "dragon": dict(width=64, height=48, spp=4, mesh=scenes.dragon_mesh(600, 12))}
for name, kw in small.items():
    ir = project.serialize_project(scenes.SCENES[name](**kw))
    o = Oracle(ir)
    with api.Renderer(0) as r:
        r.load(ir)
        assert np.array_equal(r.bvh_leaf_order(), o.bvh_leaf_order())
        res = {}
        for kind in (0, 1, 2):
            rays = o.gen_rays(kind, 50000, seed=kind)
            want, _ = o.trace(rays); got = r.trace(rays)
            same = (want["prim_id"] == got["prim_id"]) & (want["kind"] == got["kind"])
            hit = (want["kind"] != 0) & same
            rel = np.abs(want["t"][hit] - got["t"][hit]) / np.abs(want["t"][hit])
            res[f"batch{kind}"] = dict(id_mismatch=int((~same).sum()), max_rel_t=float(rel.max()) if rel.size else 0.0,
                                       t_bitexact=bool(np.array_equal(want["t"][same], got["t"][same])))
        o.render(seed=5); r.render(seed=5)
        fo, fg = o.film(), r.film()
        xo, _ = o.develop(); xg, sg = r.develop()
        res["film_weight_equal"] = bool(np.array_equal(fo[..., 1], fg[..., 1]))
        res["film_acc_max_abs"] = float(np.abs(fo[..., 0] - fg[..., 0]).max())
        res["meanY"] = [float(xo[..., 1].mean()), float(xg[..., 1].mean())]
        res["rmse_over_mean"] = float(np.sqrt(np.mean((xo[..., 1] - xg[..., 1]) ** 2)) / max(xo[..., 1].mean(), 1e-20))
        res["rays"] = [o.counters()["rays"], r.counters()["rays"]]
        out[name] = res
        print(name, res, flush=True)

# C2 timing
t0 = time.time()
mesh = scenes.dragon_mesh()
ir = project.serialize_project(scenes.dragon(spp=256, mesh=mesh))
t1 = time.time()
with api.Renderer(0) as r:
    r.load(ir)
    t2 = time.time()
    print(f"dragon: mesh+IR {t1 - t0:.1f}s, load (BVH build + upload) {t2 - t1:.1f}s, {r.info.as_dict()}", flush=True)
    timings = {}
    for spp in (1, 1, 4, 16):
        r.counters(reset=True)
        secs = r.render(seed=1, spp=spp)
        c = r.counters()
        timings[f"spp{spp}"] = dict(seconds=secs, rays=c["rays"], samples=c["path_samples"], mrays=c["rays"] / secs / 1e6,
                                   msamples=c["path_samples"] / secs / 1e6, iterations=c["wavefront_iterations"])
        print(spp, timings[f"spp{spp}"], flush=True)
    r.counters(reset=True)
    secs = r.render(seed=1, spp=1, stats=True)
    c = r.counters()
    timings["stats_spp1"] = dict(seconds=secs, rays=c["rays"], nodes_per_ray=c["nodes_visited"] / c["rays"], leaves_per_ray=c["leaves_tested"] / c["rays"])
    print(timings["stats_spp1"], flush=True)
    t3 = time.time(); xyz, srgb = r.develop(); t4 = time.time()
    timings["develop_s"] = t4 - t3
    out["dragon_c2"] = timings
    try:
        from PIL import Image
        Image.fromarray(srgb).save(ROOT / "gpurun_out" / "dragon_probe.png")
    except Exception as e:
        print("png failed", e)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "probe.json").write_text(json.dumps(out, indent=1))
print("PROBE DONE")
