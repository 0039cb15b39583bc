#!/usr/bin/env python3
"""pyr_project_load of config C5's scene (871,236 BVH items), twice per BVH builder: the first load of a process also pays for CUDA's
lazily loaded kernels and the first allocations.  PYR_BUILD_TIMING=1 prints the phases of the scene build on stderr."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pyrite_b200 import api, project, scenes

ir = project.serialize_project(scenes.SCENES[sys.argv[1] if len(sys.argv) > 1 else "bdpt_cornell_dragon"]())
for how in (sys.argv[2:] or ["gpu", "host"]):
    os.environ["PYR_BVH_BUILD"] = how
    with api.Renderer(0) as r:
        for k in range(int(os.environ.get("PYR_LOADS", "3"))):
            t = time.time(); r.load(ir); dt = time.time() - t
            print(f"{how} BVH build, load {k}: {dt:.3f} s (inside the library {r.bvh_digest()['load_seconds']:.3f} s)", flush=True)
