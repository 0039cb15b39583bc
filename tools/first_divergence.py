#!/usr/bin/env python3
"""First divergence between the product and the oracle, path sample by path sample.

For path samples (tile, i) of a scene, both sides run ONE `render_tile` iteration of the camera-to-light integrator
(pyrite/src/renderer/simple.rs:87-139 -> tracer.rs:208-345) on the same keyed RNG stream and record, per bounce: the
closest hit (kind, primitive id, t, u, v), the incident direction, the surface position, the shading normal, the scattered
direction and the Xorshift state after the bounce; plus the exposed (brightness, wavelength) pairs.  The records are
compared in path order and the FIRST field that differs is reported - with how far apart the two values are in ulps - so
that a film difference can be traced to its cause (a last-bit libm difference in a sampled direction, a flipped
visibility test, ...) instead of being allowed for.

    python tools/first_divergence.py --scene textures --samples 4000                 # product on the GPU vs the oracle
    python tools/first_divergence.py --scene textures --backend emu                  # the product's device code on the CPU (no GPU needed)
    python tools/first_divergence.py --scene textures --oracle double                # vs the oracle with double-rounded libm (pyro_math.hpp)

The product side is pyr_debug_path (C ABI) / emu_debug_path; the oracle side is pyro_debug_path.  Test infrastructure:
imports oracle/ through tests/oracle_lib.py.
"""
from __future__ import annotations

import argparse
import json
import sys
from collections import Counter
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

FIELDS = [("kind", 0, 1, "u"), ("prim_id", 1, 2, "u"), ("t", 2, 3, "f"), ("u", 3, 4, "f"), ("v", 4, 5, "f"), ("incident", 5, 8, "f"),
          ("position", 8, 11, "f"), ("normal", 11, 14, "f"), ("out", 14, 17, "f"), ("rng_w", 18, 19, "u")]


def ulps(a: np.ndarray, b: np.ndarray) -> int:
    """Largest distance in units in the last place between two float32 bit patterns (sign-magnitude -> ordered integers)."""
    def ordered(x):
        x = x.astype(np.int64)
        return np.where(x & 0x80000000, 0x80000000 - x, x)
    return int(np.abs(ordered(a) - ordered(b)).max())


def compare(rec_p: np.ndarray, rec_o: np.ndarray, exp_p: np.ndarray, exp_o: np.ndarray):
    """-> None if identical, else (bounce, field, ulps or None, product value, oracle value)."""
    for b in range(min(len(rec_p), len(rec_o))):
        p, o = rec_p[b], rec_o[b]
        last = b + 1 == len(rec_o) or b + 1 == len(rec_p)
        for name, lo, hi, ty in FIELDS:
            if name == "out" and last:
                continue  # the product only has an outgoing ray when the path goes on
            if name in ("position", "normal") and (p[19] == 0):
                continue  # an emission / miss bounce carries no surface on the product side
            if not np.array_equal(p[lo:hi], o[lo:hi]):
                view = (lambda x: x[lo:hi].view(np.float32).tolist()) if ty == "f" else (lambda x: x[lo:hi].tolist())
                return b, name, ulps(p[lo:hi], o[lo:hi]) if ty == "f" else None, view(p), view(o)
    if len(rec_p) != len(rec_o):
        return min(len(rec_p), len(rec_o)), "path_length", None, len(rec_p), len(rec_o)
    if exp_p.shape != exp_o.shape:
        return len(rec_p), "exposed_count", None, len(exp_p), len(exp_o)
    if not np.array_equal(exp_p.view(np.uint32), exp_o.view(np.uint32)):
        return len(rec_p), "brightness", ulps(exp_p.view(np.uint32).ravel(), exp_o.view(np.uint32).ravel()), exp_p[:2].tolist(), exp_o[:2].tolist()
    return None


def run(scene: str, backend: str, oracle_variant: str, samples: int, seed: int, verbose: int, overrides=None):
    from conftest import scene_ir
    from oracle_lib import Oracle

    ir = scene_ir(scene, **(overrides or {}))
    o = Oracle(ir, oracle_variant)
    if backend == "gpu":
        from pyrite_b200 import api

        prod = api.Renderer(0)
        prod.load(ir)
    else:
        from emu_lib import Emu

        prod = Emu(ir, (o.info.height, o.info.width, o.info.bins))
    tiles_x = (o.info.width + o.info.tile_size - 1) // o.info.tile_size
    tiles_y = (o.info.height + o.info.tile_size - 1) // o.info.tile_size
    n_tiles = tiles_x * tiles_y
    first = Counter()
    worst = {}
    examples = []
    diverged = 0
    rel_brightness = []
    for k in range(samples):
        tile, i = k % n_tiles, k // n_tiles
        rp, ep, _ = prod.debug_path(seed, tile, i)
        ro, eo, _ = o.debug_path(seed, tile, i)
        d = compare(rp, ro, ep, eo)
        if d is None:
            continue
        diverged += 1
        bounce, field, u, vp, vo = d
        key = field if field in ("path_length", "exposed_count", "brightness") else f"{field}@bounce{min(bounce, 3)}{'+' if bounce >= 3 else ''}"
        first[key] += 1
        if u is not None:
            worst[key] = max(worst.get(key, 0), u)
        if len(ep) and len(eo):
            denom = max(abs(float(eo[0, 0])), 1e-30)
            rel_brightness.append(abs(float(ep[0, 0]) - float(eo[0, 0])) / denom)
        if len(examples) < verbose:
            examples.append({"tile": tile, "sample": i, "bounce": bounce, "field": field, "ulps": u, "product": vp, "oracle": vo})
    rb = np.array(rel_brightness) if rel_brightness else np.zeros(0)
    report = {
        "scene": scene, "backend": backend, "oracle_libm": oracle_variant, "seed": seed, "path_samples": samples, "diverged": diverged,
        "diverged_fraction": diverged / max(samples, 1),
        "first_differing_field": dict(first.most_common()), "max_ulps_at_first_difference": worst,
        "hero_brightness_rel_diff_of_diverged": {"median": float(np.median(rb)) if rb.size else 0.0, "over_1e-3": int((rb > 1e-3).sum()), "over_5e-2": int((rb > 5e-2).sum())},
        "examples": examples,
    }
    if backend == "gpu":
        prod.close()
    return report


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--scene", default="textures", help="a scene of tests/conftest.py (small_scene_kwargs)")
    ap.add_argument("--backend", default="gpu", choices=["gpu", "emu"])
    ap.add_argument("--oracle", default="glibc", choices=["glibc", "double"], help="oracle libm variant (pyro_math.hpp)")
    ap.add_argument("--samples", type=int, default=4000)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--examples", type=int, default=3)
    ap.add_argument("--out", default=None, help="also write the report to this JSON file")
    args = ap.parse_args()
    rep = run(args.scene, args.backend, args.oracle, args.samples, args.seed, args.examples)
    text = json.dumps(rep, indent=1)
    print(text)
    if args.out:
        Path(args.out).write_text(text + "\n")


if __name__ == "__main__":
    main()
