#!/usr/bin/env python3
"""Throughput of every BASELINE config's scene on one GPU (a few sample passes each, after a warm-up pass).
Writes gpurun_out/config_rates.json."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pyrite_b200 import api, project, scenes

CONFIGS = {
    "C1 cornell 512x512 simple (64 spp)": (lambda: scenes.cornell(width=512, height=512, spp=64), 64),
    "C2 dragon 1920x1080 simple": (lambda: scenes.dragon(spp=256), 32),
    "C3 diamonds 1920x1080 simple, dispersion, S=1, B=256": (lambda: scenes.diamonds(width=1920, height=1080, spp=200), 50),
    "C4 mandelbulb + julia 3840x2160 simple": (lambda: scenes.fractals(), 4),
    "C5 textured cornell + dragon 3840x2160 bidirectional": (lambda: scenes.bdpt_cornell_dragon(), 2),
}
out = {}
for name, (make, spp) in CONFIGS.items():
    ir = project.serialize_project(make())
    with api.Renderer(0) as r:
        t0 = time.time(); r.load(ir); load_s = time.time() - t0
        r.render(seed=0, spp=1)
        r.counters(reset=True)
        secs = r.render(seed=1, spp=spp, timing=True)
        c = r.counters()
        out[name] = dict(spp=spp, seconds=secs, mrays_per_s=c["rays"] / secs / 1e6, msamples_per_s=c["path_samples"] / secs / 1e6,
                         rays_per_sample=c["rays"] / c["path_samples"], trace_s=c["trace_seconds"], shade_s=c["shade_seconds"],
                         iterations=c["wavefront_iterations"], load_s=load_s)
        print(name, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in out[name].items()}, flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "config_rates.json").write_text(json.dumps(out, indent=1))
