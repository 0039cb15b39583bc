#!/usr/bin/env python3
"""Find the path sample behind a non-finite film bin and replay it in the oracle.

A bidirectional sample whose lamp vertex has BRDF 0 exposes `x / x = 0 / 0` (the reference's `brdf_in`,
pyrite/src/renderer/bidirectional.rs:365-372), and a NaN added to a film bin stays there.  This tool shows whether a
non-finite pixel of a GPU render is such a case - i.e. whether the reference's own arithmetic produces it too:

  1. render the job on the GPU and list the non-finite film bins;
  2. for one of them, narrow the job down to ONE path sample with renders of ever smaller subsets (per-path RNG streams
     are keyed by (seed, tile, sample), so a subset reproduces its samples exactly): first the tile (pyr_render_params.
     tile_filter), then the sample index by bisection on sample_offset / sample_stride;
  3. replay exactly that sample in the oracle (pyro_render_sample) and report whether its film has the same non-finite bin.

    python tools/find_nonfinite.py --config C5 --spp 32 --seed 4242
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def nonfinite_pixels(r):
    film = r.film()
    bad = ~np.isfinite(film).all(axis=(2, 3))
    ys, xs = np.nonzero(bad)
    return list(zip(xs.tolist(), ys.tolist()))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default="C5")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--spp", type=int, default=32)
    ap.add_argument("--seed", type=int, default=4242)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import bench
    from oracle_lib import Oracle

    from pyrite_b200 import api

    ir = bench.build_project(args.config, args.small)
    report = {"config": bench.workload_name(args.config, args.small), "spp": args.spp, "seed": args.seed}
    with api.Renderer(0) as r:
        info = r.load(ir)
        r.render(seed=args.seed, spp=args.spp)
        bad = nonfinite_pixels(r)
        report["path_samples"] = int(info.width) * int(info.height) * args.spp
        report["non_finite_pixels"] = len(bad)
        print(f"{len(bad)} non-finite pixel(s) of {info.width * info.height} after {report['path_samples']} path samples: {bad[:8]}")
        if not bad:
            print(json.dumps(report, indent=1))
            return
        x, y = bad[0]
        tiles_x = (info.width + info.tile_size - 1) // info.tile_size
        tiles_y = (info.height + info.tile_size - 1) // info.tile_size
        own = (y // info.tile_size) * tiles_x + (x // info.tile_size)
        # camera-side contributions land in the sample's own tile; light-traced ones anywhere: try the own tile first, then all tiles
        tile = None
        for t in [own] + [t for t in range(tiles_x * tiles_y) if t != own]:
            r.render(seed=args.seed, spp=args.spp, only_tile=t)
            if (x, y) in nonfinite_pixels(r):
                tile = t
                break
        report["pixel"] = [x, y]
        report["tile"] = tile
        if tile is None:
            print("no single tile reproduces the pixel")
            print(json.dumps(report, indent=1))
            return
        offset, stride = 0, 1
        per_tile = info.tile_size * info.tile_size * args.spp
        while stride < per_tile:
            r.render(seed=args.seed, spp=args.spp, only_tile=tile, sample_offset=offset, sample_stride=2 * stride)
            if (x, y) not in nonfinite_pixels(r):
                offset += stride
            stride *= 2
        r.render(seed=args.seed, spp=args.spp, only_tile=tile, sample_offset=offset, sample_stride=stride)
        assert (x, y) in nonfinite_pixels(r), "bisection lost the sample"
        report["sample"] = offset
        print(f"pixel ({x}, {y}) <- tile {tile}, sample {offset}")
    o = Oracle(ir)
    o.render_sample(args.seed, tile, offset)
    film = o.film()
    obad = ~np.isfinite(film).all(axis=(2, 3))
    report["oracle_non_finite_pixels"] = [[int(a), int(b)] for b, a in zip(*np.nonzero(obad))]
    report["oracle_reproduces_it"] = bool(obad[y, x])
    print(f"the oracle's replay of that one sample: non-finite pixels {report['oracle_non_finite_pixels']} -> "
          f"{'the reference arithmetic produces it too' if obad[y, x] else 'NOT reproduced by the oracle'}")
    text = json.dumps(report, indent=1)
    print(text)
    if args.out:
        Path(args.out).write_text(text + "\n")


if __name__ == "__main__":
    main()
