import sys, os, time
sys.path.insert(0, '.')
from pyrite_b200 import api, project, scenes
seed, spp, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
ir = project.serialize_project(scenes.fractals(width=3840, height=2160))
with api.Renderer(0) as r:
    r.load(ir)
    t = time.time()
    seen = []
    r.render(seed=seed, spp=spp, stats=(mode == "stats"), timing=(mode == "timing"), progress=lambda p, m: seen.append(p) or print("progress", p, "%.2f" % (time.time() - t), flush=True) or False)
    c = r.counters()
    print(sys.argv[1:], "ok %.2f" % (time.time() - t), c["rays"], c["wavefront_iterations"], flush=True)
