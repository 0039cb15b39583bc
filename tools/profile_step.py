#!/usr/bin/env python3
"""One short render of the C2 dragon scene (default 2 spp) - the command profiled under ncu."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pyrite_b200 import api, project, scenes

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scene = sys.argv[2] if len(sys.argv) > 2 else "dragon"
kw = dict(spp=256) if scene == "dragon" else {}
if scene == "fractals":
    kw = dict(width=1920, height=1080)
if scene == "bdpt_cornell_dragon":
    kw = dict(width=1920, height=1080)
if scene == "cornell_bd":
    scene, kw = "cornell", dict(width=1024, height=1024, integrator="bidirectional")
if scene == "diamonds":
    kw = dict(width=1920, height=1080)
ir = project.serialize_project(scenes.SCENES[scene](**kw))
with api.Renderer(0) as r:
    r.load(ir)
    import os
    if not os.environ.get('PYR_NO_WARMUP'):
        r.render(seed=0, spp=1, pool_paths=int(os.environ.get('PYR_POOL', '0')))
        r.counters(reset=True)
    pool = int(os.environ.get('PYR_POOL', '0'))
    timing = os.environ.get('PYR_TIMING', '1') != '0'   # per-kernel timing runs ONE wavefront; PYR_TIMING=0 measures the production path
    secs = r.render(seed=1, spp=spp, timing=timing, pool_paths=pool)
    c = r.counters()
    if os.environ.get("PYR_COUNTS_JSON"):
        import json
        Path(os.environ["PYR_COUNTS_JSON"]).write_text(json.dumps({"scene": scene, "spp": spp, "rays": c["rays"], "path_rays": c["path_rays"], "path_samples": c["path_samples"],
                                                                   "iterations": c["wavefront_iterations"]}) + "\n")
    print(f"{scene}: {spp} spp in {secs * 1e3:.1f} ms, {c['rays'] / secs / 1e6:.0f} Mrays/s, trace {c['trace_seconds'] * 1e3:.1f} ms shade {c['shade_seconds'] * 1e3:.1f} ms, "
          f"{c['wavefront_iterations']} iterations")
