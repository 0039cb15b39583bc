#!/usr/bin/env python3
"""Replay ONE path sample of the camera-to-light integrator on the GPU (pyr_debug_path) and in the oracle (pyro_debug_path) and
print both bounce records side by side: what tools/find_nonfinite.py or first_divergence.py point at.
    python tools/replay_sample.py --config C4 --seed 4242 --tile 2534 --sample 39308"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C4")
ap.add_argument("--seed", type=int, required=True)
ap.add_argument("--tile", type=int, required=True)
ap.add_argument("--sample", type=int, required=True)
args = ap.parse_args()
import bench
from oracle_lib import Oracle

from pyrite_b200 import api

ir = bench.build_project(args.config, False)
np.set_printoptions(precision=7, linewidth=220)
o = Oracle(ir)
ro, eo, po = o.debug_path(args.seed, args.tile, args.sample)
with api.Renderer(0) as r:
    r.load(ir)
    rg, eg, pg = r.debug_path(args.seed, args.tile, args.sample)
for name, rec, exp in (("gpu", rg, eg), ("oracle", ro, eo)):
    print(f"--- {name}: {len(rec)} bounces, exposed {exp[:3].tolist()}")
    for b, x in enumerate(rec):
        f = x[2:17].view(np.float32)
        print(f"  {b}: kind {x[0]} id {x[1]} t {f[0]:.7g} incident {f[3:6]} position {f[6:9]} normal {f[9:12]} out {f[12:15]} rays {x[17]} rng {x[18]:08x}")
