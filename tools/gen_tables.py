#!/usr/bin/env python3
"""Bake the colour-science constant tables the render path needs into pyrite_b200/data/tables.npz.

The reference bakes the same four public data sets into `const` arrays at build time
(pyrite/build.rs:18-59 Burns sRGB basis, :68-121 CIE 1931 2-degree observer at 1 nm,
:131-187 CIE illuminants D65 and A at 5 nm; note `A` is divided by 100, `D65` is not,
build.rs:148,157, and the Burns table's max is 360 + row count = 831, build.rs:37-38).
They are numeric data, not code; this script only re-serialises them.  It runs in the
build container (where /root/reference exists); the .npz it writes is committed so that
nothing reads /root/reference at run time.
"""
import csv
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/pyrite/data")
OUT = Path(__file__).resolve().parent.parent / "pyrite_b200" / "data" / "tables.npz"


def rows(name):
    with open(REF / name, newline="") as f:
        return list(csv.DictReader(f))


def main():
    burns = rows("srgb_cie1931.csv")
    xyz = rows("ciexyz65_1.csv")
    d65 = rows("d65.csv")
    a = rows("a.csv")
    f32 = np.float32
    out = {
        # Spectrum::Array{min:360, max:360+len} (build.rs:37-38)
        "burns_min": f32(360.0),
        "burns_max": f32(360.0 + len(burns)),
        "burns_rgb": np.array([[f32(r["r"]), f32(r["g"]), f32(r["b"])] for r in burns], dtype=f32),
        "xyz_min": f32(min(f32(r["wavelength"]) for r in xyz)),
        "xyz_max": f32(max(f32(r["wavelength"]) for r in xyz)),
        "xyz": np.array([[f32(r["x"]), f32(r["y"]), f32(r["z"])] for r in xyz], dtype=f32),
        # min/max are taken over BOTH illuminant files (build.rs:135-160)
        "illum_min": f32(min(min(f32(r["wavelength"]) for r in d65), min(f32(r["wavelength"]) for r in a))),
        "illum_max": f32(max(max(f32(r["wavelength"]) for r in d65), max(f32(r["wavelength"]) for r in a))),
        "d65": np.array([f32(r["intensity"]) for r in d65], dtype=f32),
        # f32 parse then f32 divide, as `intensity / 100.0` on an f32 field (build.rs:157)
        "a": np.array([f32(r["intensity"]) / f32(100.0) for r in a], dtype=f32),
    }
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", v))


if __name__ == "__main__":
    main()
