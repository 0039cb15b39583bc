"""`python -m pyrite_b200 path/to/project.lua` - the command-line driver of pyrite (pyrite/src/main.rs:52-332)
on the B200 path: load the project (Lua), build the scene, render with progress, develop, write `render.png`
next to the project file.  A preview is written at most every 20 s with 30 nm integration steps
(main.rs:261-299); timings are printed like main.rs:90-102."""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="pyrite_b200", description=__doc__)
    ap.add_argument("project", help="path to a pyrite project.lua")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--seed", type=int, default=None, help="default: from the clock (the reference seeds from the OS)")
    ap.add_argument("--spp", type=int, default=0, help="override renderer.pixel_samples")
    ap.add_argument("--out", default=None, help="output image (default: render.png next to the project file, main.rs:180-184)")
    ap.add_argument("--no-preview", action="store_true")
    ap.add_argument("--preview-interval", type=float, default=20.0, help="seconds between preview images (main.rs:261: 20)")
    ap.add_argument("--pool", type=int, default=0, help="paths in flight (0 = library default); progress is reported between wavefront batches")
    args = ap.parse_args(argv)

    from PIL import Image

    from . import api, lua_project

    total_start = time.time()
    path = Path(args.project)
    try:
        ir = lua_project.load_project_ir(path)
    except (lua_project.ProjectLoadError, OSError) as e:
        print(f"error while loading project file: {e}", file=sys.stderr)   # main.rs:68-71
        return 1
    out = Path(args.out) if args.out else path.resolve().parent / "render.png"
    try:
        r = api.Renderer(args.device)
        info = r.load(ir)
    except api.PyriteError as e:
        print(f"error while parsing project: {e}", file=sys.stderr)        # main.rs:104
        return 1
    loaded = time.time()
    print(f"Project loading: {loaded - total_start:.3f} s  ({info.n_objects} objects, {info.n_bvh_nodes} BVH nodes, "
          f"{info.width}x{info.height}, {'bidirectional' if info.algorithm else 'simple'} x {args.spp or info.pixel_samples} spp)")
    state = {"last": time.time(), "pct": -1, "previews": 0}

    def progress(pct, message):
        if pct != state["pct"]:
            state["pct"] = pct
            print(f"\r{message}: {pct:3d}%", end="", flush=True)
        if not args.no_preview and time.time() - state["last"] >= args.preview_interval and pct < 100:
            _, srgb = r.develop(30.0, want_xyz=False)   # the preview integrates in 30 nm steps (main.rs:274-279)
            Image.fromarray(srgb).save(out)
            state["last"] = time.time()
            state["previews"] += 1
        return False

    seed = args.seed if args.seed is not None else time.time_ns() & 0xFFFFFFFFFFFF
    device_seconds = r.render(seed=seed, spp=args.spp, pool_paths=args.pool, progress=progress)
    rendered = time.time()
    print()
    _, srgb = r.develop(2.0, want_xyz=False)
    Image.fromarray(srgb).save(out)
    c = r.counters()
    print(f"Rendering: {rendered - loaded:.3f} s  (device {device_seconds:.3f} s, {c['rays'] / max(device_seconds, 1e-9) / 1e6:.0f} Mrays/s, "
          f"{c['path_samples'] / max(device_seconds, 1e-9) / 1e6:.1f} M path samples/s)")
    print(f"Total: {time.time() - total_start:.3f} s; wrote {out} ({state['previews']} previews on the way)")
    r.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
