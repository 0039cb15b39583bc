#!/usr/bin/env python3
"""Generate the committed input fixtures under tests/golden/ from the reference's example assets.

Runs only in the build container (reads /root/reference); the outputs are committed so that
the GPU box, which has no /root/reference, can run the parity tests.

  meshes/*.npz    the reference's OBJ meshes (pyrite/test/*/*.obj) parsed with
                  pyrite_b200.project.parse_obj and stored as index arrays
  spectra.json    the array spectra of pyrite/test/cornell/{colors,lamp}.lua (numbers only)
  textures/*.png  down-scaled copies of the JPG textures used by pyrite/test/textures
"""
import json
import re
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from pyrite_b200.project import load_obj  # noqa: E402

REF = Path("/root/reference/pyrite/test")
OUT = Path(__file__).resolve().parent


def lua_spectrum(text, name):
    m = re.search(name + r"\s*=\s*spectrum\s*\{(.*?)points\s*=\s*\{(.*?)\}", text, re.S)
    head, pts = m.group(1), m.group(2)
    lo = float(re.search(r"min\s*=\s*([-0-9.eE]+)", head).group(1))
    hi = float(re.search(r"max\s*=\s*([-0-9.eE]+)", head).group(1))
    points = [float(x) for x in re.findall(r"[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?", pts)]
    return {"format": "array", "min": lo, "max": hi, "points": points}


def main():
    (OUT / "meshes").mkdir(exist_ok=True)
    for rel in ["cornell/box.obj", "diamonds/diamonds.obj", "snowflake/snowflake.obj", "textures/cube.obj", "textures/color_checker.obj"]:
        mesh = load_obj(REF / rel)
        dst = OUT / "meshes" / (Path(rel).stem + ".npz")
        mesh.save(dst)
        print(dst.name, [(n, len(t)) for n, t in mesh.objects])
    colors = (REF / "cornell/colors.lua").read_text()
    lamp = (REF / "cornell/lamp.lua").read_text()
    spectra = {k: lua_spectrum(colors, k) for k in ("white", "green", "red")}
    spectra["lamp"] = lua_spectrum(lamp, "color")
    (OUT / "spectra.json").write_text(json.dumps(spectra))
    print({k: len(v["points"]) for k, v in spectra.items()})
    try:
        from PIL import Image
        (OUT / "textures").mkdir(exist_ok=True)
        for rel, size in [("textures/tiles/color.jpg", 256), ("textures/tiles/normal.jpg", 256), ("textures/color_checker.jpg", 256),
                          ("textures/tactile_paving/color.jpg", 128), ("textures/tactile_paving/normal.jpg", 128)]:
            im = Image.open(REF / rel).convert("RGB")
            im = im.resize((size, max(1, size * im.height // im.width)), Image.BILINEAR)
            name = rel.replace("textures/", "").replace("/", "_").replace(".jpg", ".png")
            im.save(OUT / "textures" / name, optimize=True)
            print(name, im.size)
    except Exception as e:  # pragma: no cover
        print("texture fixtures skipped:", e)


if __name__ == "__main__":
    main()
