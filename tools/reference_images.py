#!/usr/bin/env python3
"""SURVEY.md §8(c) pin (3): renders of the reference's own, unmodified example projects compared with the example images
the reference ships (pyrite/test/*/hq_example.png) - the only reference-held outputs that exist for this path.

    python tools/reference_images.py --make-fixtures      # here, where /root/reference exists: writes tests/golden/reference_scenes/
    python tools/reference_images.py --backend gpu        # on the B200 box (or --backend oracle on the CPU, lower spp)

Fixtures (data, not code): the project IR of each unmodified `*.lua` as loaded by pyrite_b200.lua_project (what
`load_project` hands to `parse_project`, main.rs:111-134) and the example PNG reduced to 8x8-block mean sRGB.

What can and cannot be concluded.  The example images were made by OLDER versions of pyrite and of the scene files
(cornell/hq_example.png has no fractal although cornell.lua has one; snowflake's background is ~20x brighter than the
current scene's lamps (`d65 * 6` with d65.csv normalised to 1 at 560 nm) can make it; dragon.lua still carries spot-light keys
the current point light ignores), so they pin FRAMING and STRUCTURE - camera convention, which object is
where, silhouettes, which side is lit - not radiometry.  The metric is therefore the Pearson correlation of log block
luminance, reported next to the same number for the horizontally mirrored render (what a flipped camera convention would
give) and the mean-luminance ratio.
"""
from __future__ import annotations

import argparse
import gzip
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
FIX = ROOT / "tests" / "golden" / "reference_scenes"
REF = Path("/root/reference/pyrite/test")
SCENES = ["spheres", "diamonds", "colors", "snowflake", "cornell"]
BLOCK = 8


def block_means(img: np.ndarray, b: int = BLOCK) -> np.ndarray:
    h, w, c = img.shape
    return img[: h // b * b, : w // b * b].astype(np.float64).reshape(h // b, b, w // b, b, c).mean(axis=(1, 3))


def srgb_to_linear(v):
    v = np.asarray(v, np.float64) / 255.0
    return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)


def luminance(blocks_srgb):
    lin = srgb_to_linear(blocks_srgb)
    return 0.2126 * lin[..., 0] + 0.7152 * lin[..., 1] + 0.0722 * lin[..., 2]


def pearson(a, b):
    a, b = a.ravel() - a.mean(), b.ravel() - b.mean()
    return float((a * b).sum() / max(np.sqrt((a * a).sum() * (b * b).sum()), 1e-30))


def make_fixtures():
    from PIL import Image

    from pyrite_b200 import lua_project

    FIX.mkdir(parents=True, exist_ok=True)
    blocks = {}
    for name in SCENES:
        ir = lua_project.load_project_ir(REF / name / f"{name}.lua")
        (FIX / f"{name}.ir.gz").write_bytes(gzip.compress(ir, 9, mtime=0))
        img = np.asarray(Image.open(REF / name / "hq_example.png").convert("RGB"))
        blocks[name] = np.round(block_means(img)).astype(np.uint8)
        print(f"{name}: IR {len(ir)} bytes, example image {img.shape[1]}x{img.shape[0]} -> {blocks[name].shape[1]}x{blocks[name].shape[0]} blocks")
    np.savez_compressed(FIX / "example_image_blocks.npz", **blocks)


def load_ir(name: str) -> bytes:
    return gzip.decompress((FIX / f"{name}.ir.gz").read_bytes())


def compare(name: str, srgb: np.ndarray) -> dict:
    ref = np.load(FIX / "example_image_blocks.npz")[name].astype(np.float64)
    mine = block_means(srgb)
    lr, lm = np.log(luminance(ref) + 1e-3), np.log(luminance(mine) + 1e-3)
    return {
        "pearson_log_luminance": pearson(lr, lm),
        "pearson_if_mirrored": pearson(lr, lm[:, ::-1]),
        "mean_luminance_ratio": float(luminance(mine).mean() / max(luminance(ref).mean(), 1e-9)),
        "mean_abs_block_srgb_diff": float(np.abs(ref - mine).mean()),
    }


def render(name: str, backend: str, spp: int, seed: int = 1):
    ir = load_ir(name)
    if backend == "gpu":
        from pyrite_b200 import api

        with api.Renderer(0) as r:
            r.load(ir)
            r.render(seed=seed, spp=spp)
            return r.develop(want_xyz=False)[1]
    from oracle_lib import Oracle

    o = Oracle(ir)
    o.render(seed=seed, spp=spp)
    return o.develop()[1]


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--make-fixtures", action="store_true")
    ap.add_argument("--backend", default="gpu", choices=["gpu", "oracle"])
    ap.add_argument("--spp", type=int, default=0, help="0 = each project's own pixel_samples")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if args.make_fixtures:
        make_fixtures()
        return
    table = {name: compare(name, render(name, args.backend, args.spp)) for name in SCENES}
    text = json.dumps({"backend": args.backend, "spp": args.spp or "project", "scenes": table}, indent=1)
    print(text)
    if args.out:
        Path(args.out).write_text(text + "\n")


if __name__ == "__main__":
    main()
