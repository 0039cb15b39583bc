import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import scene_ir
from oracle_lib import Oracle
from pyrite_b200 import api
ir = scene_ir("fractals")
o = Oracle(ir)
np.set_printoptions(precision=7, linewidth=220)
with api.Renderer(0) as r:
    info = r.load(ir)
    r.render(seed=4242, spp=8)
    film = r.film()
    bad = ~np.isfinite(film).all(axis=(2, 3))
    ys, xs = np.nonzero(bad)
    print("non-finite pixels", list(zip(xs.tolist(), ys.tolist())))
    tiles_x = (info.width + info.tile_size - 1) // info.tile_size
    for x, y in list(zip(xs.tolist(), ys.tolist()))[:2]:
        tile = (y // info.tile_size) * tiles_x + (x // info.tile_size)
        per_tile = info.tile_size * info.tile_size * 8
        found = None
        for i in range(per_tile):
            rg, eg, pg = r.debug_path(4242, tile, i)
            if not np.isfinite(eg).all():
                found = i
                break
        print("pixel", x, y, "tile", tile, "sample", found)
        if found is None:
            continue
        ro, eo, po = o.debug_path(4242, tile, found)
        for name, rec, exp in (("gpu", rg, eg), ("oracle", ro, eo)):
            print(f"--- {name}: {len(rec)} bounces, exposed {exp[:2].tolist()}")
            for b, q in enumerate(rec):
                f = q[2:17].view(np.float32)
                print(f"  {b}: kind {q[0]} id {q[1]} t {f[0]:.7g} incident {f[3:6]} position {f[6:9]} normal {f[9:12]} out {f[12:15]} rays {q[17]} rng {q[18]:08x}")
