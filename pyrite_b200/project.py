"""Project surface: a Python mirror of pyrite's Lua project DSL, serialised to the "project IR".

The reference evaluates a `project.lua` with an embedded Lua plus `project/lib.lua`
(pyrite/src/project/lib.lua:1-309) and decodes the returned table into typed structs
(`project::Project`, pyrite/src/project/mod.rs:103-269).  Everything *after* that point
(`parse_project`, main.rs:111-134: Camera/Renderer/World::from_project, program compile,
BVH build, render) is done natively behind the C ABI (include/pyrite_b200.h).

This module provides the same builder vocabulary as lib.lua (`shape.sphere{...}`,
`material.diffuse{...}`, `mix`, `fresnel`, `vector`, `rgb`, `spectrum`, `texture`,
`light_source.d65`, `transform.look_at`, `camera.perspective`, `renderer.simple`, ...) so
that a scene reads like the Lua original, and `serialize_project` which performs the
typed decode (same field names, same optionality, unknown keys ignored - SURVEY.md §9 Q11)
and writes the flat little-endian IR that `pyr_project_load` consumes.

IR layout (version 1) - see DESIGN.md §3 for the authoritative description.
"""
from __future__ import annotations

import math
import struct
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np

IR_MAGIC = 0x52495950  # 'PYIR'
IR_VERSION = 1

_DATA = Path(__file__).resolve().parent / "data"


# --------------------------------------------------------------------------- DSL nodes
class Node(dict):
    """A Lua-table stand-in: a dict with identity (lib.lua `assign_id`, :12-15) and the
    `basics_mt` helpers `clone`/`with` (lib.lua:44-73)."""

    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def clone(self):
        return type(self)(self)

    def with_(self, new=None, **kw):
        c = self.clone()
        if callable(new):
            new = new(c)
        if new:
            c.update(new)
        c.update(kw)
        return c


class Expr(Node):
    """Tables created through `_pyrite.make_expression` (lib.lua:79-82): arithmetic
    metamethods build `binary` nodes (lib.lua:2-11, 83-97)."""

    def _bin(op, swap=False):  # noqa: N805
        def f(self, other):
            lhs, rhs = (other, self) if swap else (self, other)
            return Expr(type="binary", operator=op, lhs=lhs, rhs=rhs)

        return f

    __add__ = _bin("add")
    __radd__ = _bin("add", True)
    __sub__ = _bin("sub")
    __rsub__ = _bin("sub", True)
    __mul__ = _bin("mul")
    __rmul__ = _bin("mul", True)
    __truediv__ = _bin("div")
    __rtruediv__ = _bin("div", True)
    del _bin

    def mix(self, other, amount):
        return mix(self, other, amount)


def mix(lhs, rhs=None, amount=None):
    """lib.lua:99-112"""
    if isinstance(lhs, dict) and "type" not in lhs and rhs is None:
        return Expr(lhs, type="mix")
    return Expr(type="mix", lhs=lhs, rhs=rhs, amount=amount)


def fresnel(ior, env_ior=None):
    """lib.lua:114-119"""
    return Expr(type="fresnel", ior=ior, env_ior=1 if env_ior is None else env_ior)


def vector(x=None, y=None, z=None, w=None, **kw):
    """lib.lua:122-145; `vector{z=1}` == vector(z=1)."""
    if isinstance(x, dict) and "type" not in x:
        kw, x = x, None
    if kw:
        return Expr(type="vector", x=kw.get("x", 0.0), y=kw.get("y", 0.0), z=kw.get("z", 0.0), w=kw.get("w", 0.0))
    return Expr(type="vector", x=x or 0.0, y=y or 0.0, z=z or 0.0, w=w or 0.0)


def blackbody(temperature):
    """lib.lua:147-152"""
    return Expr(type="blackbody", temperature=temperature)


def spectrum(properties=None, **kw):
    """lib.lua:154-159"""
    p = dict(properties or {})
    p.update(kw)
    p["type"] = "spectrum"
    return Expr(p)


def rgb(red=None, green=None, blue=None):
    """lib.lua:161-171"""
    return Expr(type="rgb", red=red or 0.0, green=green or 0.0, blue=blue or 0.0)


def texture(path, *modifiers):
    """lib.lua:173-190.  `path` may also be an in-memory `Image` (synthetic textures)."""
    p = Expr(type="color_texture", path=path, linear=False, mono=False)
    for m in modifiers:
        if isinstance(m, str):
            p[m] = True
    if p["mono"]:
        p["type"] = "mono_texture"
    return p


def _maker(type_name, cls=Node):
    def make(properties=None, **kw):
        p = cls(properties or {})
        p.update(kw)
        p["type"] = type_name
        return p

    return make


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


shape = _NS(sphere=_maker("sphere"), plane=_maker("plane"), mesh=_maker("mesh"), ray_marched=_maker("ray_marched"))
ray_marched = _NS(quaternion_julia=_maker("quaternion_julia"), mandelbulb=_maker("mandelbulb"))
# lib.lua:228-230 defines only `cubic`; the Rust side also accepts "regular"/"bicomplex" (world.rs:158-161)
quaternion_julia = _NS(
    cubic=Node(type="quaternion_julia", name="cubic"),
    regular=Node(type="quaternion_julia", name="regular"),
    bicomplex=Node(type="quaternion_julia", name="bicomplex"),
)
bounds = _NS(box=_maker("box"), sphere=_maker("sphere"))  # `sphere` accepted by mod.rs:211-214
material = _NS(
    diffuse=_maker("diffuse", Expr),
    emissive=_maker("emissive", Expr),
    mirror=_maker("mirror", Expr),
    refractive=_maker("refractive", Expr),
)
light_source = _NS(d65=Expr(type="spectrum", name="d65"), a=Expr(type="spectrum", name="a"))
transform = _NS(look_at=_maker("look_at"))
camera = _NS(perspective=_maker("perspective"))
renderer = _NS(simple=_maker("simple"), bidirectional=_maker("bidirectional"), photon_mapping=_maker("photon_mapping"))
light = _NS(point=_maker("point_light"), directional=_maker("directional_light"))


# --------------------------------------------------------------------------- assets
class Mesh:
    """In-memory equivalent of `obj::ObjData` after `remove_materials`
    (pyrite/src/project/meshes.rs:63-107): shared position / texture / normal pools and,
    per object, the 3-vertex polygons as (v, t, n) index triples with -1 for `None`.
    Non-triangle polygons are dropped here, as `World::from_project` does (world.rs:218-232).
    """

    def __init__(self, position, texture, normal, objects):
        self.position = np.ascontiguousarray(position, dtype=np.float32).reshape(-1, 3)
        self.texture = np.ascontiguousarray(texture, dtype=np.float32).reshape(-1, 2)
        self.normal = np.ascontiguousarray(normal, dtype=np.float32).reshape(-1, 3)
        # list of (name, int32 array [n_tris, 3, 3])
        self.objects = [(n, np.ascontiguousarray(t, dtype=np.int32).reshape(-1, 3, 3)) for n, t in objects]

    @property
    def n_triangles(self):
        return sum(len(t) for _, t in self.objects)

    def save(self, path):
        d = {"position": self.position, "texture": self.texture, "normal": self.normal,
             "names": np.array([n for n, _ in self.objects])}
        for i, (_, t) in enumerate(self.objects):
            d[f"tris_{i}"] = t
        np.savez_compressed(path, **d)

    @staticmethod
    def load(path):
        z = np.load(path)
        names = [str(n) for n in z["names"]]
        return Mesh(z["position"], z["texture"], z["normal"], [(n, z[f"tris_{i}"]) for i, n in enumerate(names)])


def parse_obj(text: str) -> Mesh:
    """Wavefront OBJ reader with the `obj` 0.10.2 crate's grouping rules:
    `o name` starts an object, `g name` a group inside it, faces before any `o` go to an
    object named "default"; indices are 1-based, negative = relative; `v`, `v/vt`,
    `v//vn`, `v/vt/vn`.  Material statements are ignored (meshes.rs:54-57)."""
    pos: List[List[float]] = []
    tex: List[List[float]] = []
    nor: List[List[float]] = []
    objects: List[tuple] = []
    cur_name = "default"
    cur: List[List[int]] = []
    started = False

    def flush():
        nonlocal cur
        if cur or started:
            if cur:
                objects.append((cur_name, np.array(cur, dtype=np.int32).reshape(-1, 3, 3)))
        cur = []

    def idx(s, n):
        i = int(s)
        return i - 1 if i > 0 else n + i

    for line in text.splitlines():
        line = line.split("#", 1)[0].strip()
        if not line:
            continue
        parts = line.split()
        k = parts[0]
        if k == "v":
            pos.append([float(x) for x in parts[1:4]])
        elif k == "vt":
            uv = [float(x) for x in parts[1:3]]
            tex.append(uv + [0.0] * (2 - len(uv)))
        elif k == "vn":
            nor.append([float(x) for x in parts[1:4]])
        elif k == "o":
            flush()
            cur_name = parts[1] if len(parts) > 1 else ""
            started = True
        elif k == "f":
            tuples = []
            for p in parts[1:]:
                f = p.split("/")
                v = idx(f[0], len(pos))
                t = idx(f[1], len(tex)) if len(f) > 1 and f[1] else -1
                n = idx(f[2], len(nor)) if len(f) > 2 and f[2] else -1
                tuples.append((v, t, n))
            if len(tuples) == 3:  # world.rs:218-232 keeps only `[x, y, z]` polygons
                cur.append([list(t) for t in tuples])
    flush()
    return Mesh(pos or np.zeros((0, 3)), tex or np.zeros((0, 2)), nor or np.zeros((0, 3)), objects)


def load_obj(path) -> Mesh:
    return parse_obj(Path(path).read_text())


def srgb_to_linear(v: np.ndarray) -> np.ndarray:
    """sRGB EOTF, as palette's `Srgb::into_linear` (texture.rs:218-224)."""
    v = v.astype(np.float64)
    return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4).astype(np.float32)


class Image:
    """Decoded 8-bit image (H, W, C) with C in {1, 2, 3, 4}; stands in for `image::DynamicImage`."""

    def __init__(self, pixels: np.ndarray):
        px = np.asarray(pixels)
        if px.ndim == 2:
            px = px[:, :, None]
        assert px.dtype == np.uint8 and px.ndim == 3 and px.shape[2] in (1, 2, 3, 4)
        self.pixels = px

    @staticmethod
    def open(path):
        from PIL import Image as PILImage

        im = PILImage.open(path)
        if im.mode not in ("L", "LA", "RGB", "RGBA"):
            im = im.convert("RGB")
        return Image(np.asarray(im))

    def linear_floats(self, linear: bool) -> np.ndarray:
        """`convert_pixels` (texture.rs:175-198): components / 255, colour channels through the
        sRGB EOTF unless `linear`; alpha is never gamma-decoded (texture.rs:289-300)."""
        px = self.pixels.astype(np.float32) / np.float32(255.0)
        c = px.shape[2]
        ncol = c if c in (1, 3) else c - 1
        if not linear:
            px = px.copy()
            px[:, :, :ncol] = srgb_to_linear(px[:, :, :ncol])
        return px

    def to_lin_srgba(self, linear: bool) -> np.ndarray:
        """`Texture<LinSrgba>` (FromColor: luma -> r=g=b, missing alpha -> 1)."""
        px = self.linear_floats(linear)
        h, w, c = px.shape
        out = np.ones((h, w, 4), dtype=np.float32)
        if c in (1, 2):
            out[:, :, 0:3] = px[:, :, 0:1]
            if c == 2:
                out[:, :, 3] = px[:, :, 1]
        else:
            out[:, :, 0:3] = px[:, :, 0:3]
            if c == 4:
                out[:, :, 3] = px[:, :, 3]
        return out

    def to_lin_luma(self, linear: bool) -> np.ndarray:
        """`Texture<LinLuma>`: luma = Y row of the sRGB->XYZ matrix (SURVEY.md §10 palette)."""
        px = self.linear_floats(linear)
        c = px.shape[2]
        if c in (1, 2):
            return np.ascontiguousarray(px[:, :, 0])
        r, g, b = px[:, :, 0], px[:, :, 1], px[:, :, 2]
        return (np.float32(0.2126729) * r + np.float32(0.7151522) * g + np.float32(0.0721750) * b).astype(np.float32)


def builtin_tables() -> Dict[str, np.ndarray]:
    z = np.load(_DATA / "tables.npz")
    return {k: z[k] for k in z.files}


# --------------------------------------------------------------------------- IR writer
class _W:
    def __init__(self):
        self.parts: List[bytes] = []

    def u32(self, v):
        self.parts.append(struct.pack("<I", int(v) & 0xFFFFFFFF))

    def f32(self, v):
        self.parts.append(struct.pack("<f", float(v)))

    def f64(self, v):
        self.parts.append(struct.pack("<d", float(v)))

    def string(self, s: str):
        b = s.encode("utf-8")
        self.u32(len(b))
        self.parts.append(b + b"\0" * ((4 - len(b) % 4) % 4))

    def array(self, a: np.ndarray, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        self.parts.append(a.tobytes())

    def opt_u32(self, v):
        if v is None:
            self.u32(0)
            self.u32(0)
        else:
            self.u32(1)
            self.u32(v)

    def bytes(self) -> bytes:
        return b"".join(self.parts)


_EXPR_TYPES = {"vector": 0, "rgb": 1, "binary": 2, "mix": 3, "clamp": 4, "fresnel": 5, "blackbody": 6,
               "spectrum": 7, "color_texture": 8, "mono_texture": 9}
_BIN_OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3}
_MAT_TYPES = {"emissive": 0, "diffuse": 1, "mirror": 2, "refractive": 3, "mix": 4, "binary": 5}


class ProjectError(ValueError):
    pass


class _Serializer:
    """The typed decode of `typed_nodes::FromLua`: nodes are interned by
    table identity, unknown keys are ignored, missing required keys are errors."""

    def __init__(self, base_dir: Optional[Path]):
        self.base_dir = Path(base_dir) if base_dir else None
        self.expr_ids: Dict[int, int] = {}
        self.expr_recs: List[Any] = []
        self.mat_ids: Dict[int, int] = {}
        self.mat_recs: List[Any] = []
        self.spectrum_ids: Dict[int, int] = {}
        self.spectra: List[Any] = []
        self.color_tex_ids: Dict[Any, int] = {}
        self.color_tex: List[np.ndarray] = []
        self.mono_tex_ids: Dict[Any, int] = {}
        self.mono_tex: List[np.ndarray] = []
        self.mesh_ids: Dict[Any, int] = {}
        self.meshes: List[Mesh] = []
        self._keep: List[Any] = []  # keep nodes alive so id() stays unique
        self.tables = builtin_tables()

    # -- expressions
    def expr(self, v):
        if isinstance(v, bool):
            raise ProjectError("expected a number or an expression table, found a boolean")
        if isinstance(v, (int, float, np.integer, np.floating)):
            return (0, float(v))
        if isinstance(v, dict):
            return (1, self.expr_node(v))
        raise ProjectError(f"expected a number or an expression table, found {type(v).__name__}")

    def need(self, t, k):
        if k not in t or t[k] is None:
            raise ProjectError(f"missing field '{k}' in {t.get('type', 'table')}")
        return t[k]

    def expr_node(self, t) -> int:
        key = id(t)
        if key in self.expr_ids:
            return self.expr_ids[key]
        self._keep.append(t)
        ty = t.get("type")
        if ty not in _EXPR_TYPES:
            raise ProjectError(f"unknown expression type: {ty!r}")
        idx = len(self.expr_recs)
        self.expr_ids[key] = idx
        self.expr_recs.append(None)
        code = _EXPR_TYPES[ty]
        if ty == "vector":
            rec = (code, [self.expr(self.need(t, k)) for k in "xyzw"])
        elif ty == "rgb":
            rec = (code, [self.expr(self.need(t, k)) for k in ("red", "green", "blue")])
        elif ty == "binary":
            op = self.need(t, "operator")
            if op not in _BIN_OPS:
                raise ProjectError(f"unknown binary operator: {op!r}")
            rec = (code, _BIN_OPS[op], [self.expr(self.need(t, "lhs")), self.expr(self.need(t, "rhs"))])
        elif ty == "mix":
            rec = (code, [self.expr(self.need(t, k)) for k in ("amount", "lhs", "rhs")])
        elif ty == "clamp":
            rec = (code, [self.expr(self.need(t, k)) for k in ("value", "min", "max")])
        elif ty == "fresnel":
            rec = (code, [self.expr(self.need(t, "ior")), self.expr(self.need(t, "env_ior"))])
        elif ty == "blackbody":
            rec = (code, [self.expr(self.need(t, "temperature"))])
        elif ty == "spectrum":
            rec = (code, self.spectrum(t))
        elif ty == "color_texture":
            rec = (code, self.color_texture(t))
        else:
            rec = (code, self.mono_texture(t))
        self.expr_recs[idx] = rec
        return idx

    def spectrum(self, t) -> int:
        """`SpectrumId::from_lua` (project/spectra.rs:120-145)."""
        key = id(t)
        if key in self.spectrum_ids:
            return self.spectrum_ids[key]
        self._keep.append(t)
        if isinstance(t.get("name"), str):
            name = t["name"]
            if name not in ("a", "d65"):
                raise ProjectError(f"unknown builtin spectrum: {name}")
            rec = ("array", float(self.tables["illum_min"]), float(self.tables["illum_max"]), self.tables[name])
        else:
            fmt = t.get("format")
            if fmt == "array":
                pts = np.array([float(p) for p in self.need(t, "points")], dtype=np.float32)
                rec = ("array", float(self.need(t, "min")), float(self.need(t, "max")), pts)
            elif fmt == "curve":
                pts = np.array([[float(p[0]), float(p[1])] for p in self.need(t, "points")], dtype=np.float32).reshape(-1, 2)
                rec = ("curve", pts)
            else:
                raise ProjectError(f"unknown spectrum format: {fmt!r}")
        self.spectrum_ids[key] = len(self.spectra)
        self.spectra.append(rec)
        return self.spectrum_ids[key]

    def _image(self, path) -> tuple:
        if isinstance(path, Image):
            return ("mem", id(path)), path
        p = Path(path)
        if self.base_dir and not p.is_absolute():
            p = self.base_dir / p
        p = p.resolve()
        if not p.exists():
            raise ProjectError(f"could not load {p} as texture: file not found")
        return ("file", str(p)), None

    def color_texture(self, t) -> int:
        """`TextureLoader::load_color` (project/textures.rs:68-89): one texture per file."""
        key, img = self._image(self.need(t, "path"))
        if key not in self.color_tex_ids:
            img = img or Image.open(key[1])
            self._keep.append(img)
            self.color_tex_ids[key] = len(self.color_tex)
            self.color_tex.append(img.to_lin_srgba(bool(t.get("linear", False))))
        return self.color_tex_ids[key]

    def mono_texture(self, t) -> int:
        key, img = self._image(self.need(t, "path"))
        if key not in self.mono_tex_ids:
            img = img or Image.open(key[1])
            self._keep.append(img)
            self.mono_tex_ids[key] = len(self.mono_tex)
            self.mono_tex.append(img.to_lin_luma(bool(t.get("linear", False))))
        return self.mono_tex_ids[key]

    # -- materials
    def mat_node(self, t) -> int:
        """`SurfaceMaterial` (project/materials.rs:7-34)."""
        if not isinstance(t, dict):
            raise ProjectError("expected a surface material table")
        key = id(t)
        if key in self.mat_ids:
            return self.mat_ids[key]
        self._keep.append(t)
        ty = t.get("type")
        if ty not in _MAT_TYPES:
            raise ProjectError(f"unknown surface material type: {ty!r}")
        idx = len(self.mat_recs)
        self.mat_ids[key] = idx
        self.mat_recs.append(None)
        code = _MAT_TYPES[ty]
        if ty in ("emissive", "diffuse", "mirror"):
            rec = (code, self.expr(self.need(t, "color")))
        elif ty == "refractive":
            rec = (code, self.expr(self.need(t, "color")), self.expr(self.need(t, "ior")),
                   [None if t.get(k) is None else self.expr(t[k]) for k in ("dispersion", "env_ior", "env_dispersion")])
        elif ty == "mix":
            rec = (code, self.mat_node(self.need(t, "lhs")), self.mat_node(self.need(t, "rhs")), self.expr(self.need(t, "amount")))
        else:
            if self.need(t, "operator") != "add":
                raise ProjectError("materials only support the `add` operator")
            rec = (code, self.mat_node(self.need(t, "lhs")), self.mat_node(self.need(t, "rhs")))
        self.mat_recs[idx] = rec
        return idx

    def material(self, t):
        """`project::Material{surface, normal_map}` (project/mod.rs:240-244)."""
        if not isinstance(t, dict):
            raise ProjectError("expected a material table")
        nm = t.get("normal_map")
        return (self.mat_node(self.need(t, "surface")), None if nm is None else self.expr(nm))

    def mesh(self, f) -> int:
        if isinstance(f, Mesh):
            key = ("mem", id(f))
            m = f
        else:
            p = Path(f)
            if self.base_dir and not p.is_absolute():
                p = self.base_dir / p
            p = p.resolve()
            key = ("file", str(p))
            m = None
        if key not in self.mesh_ids:
            if m is None:
                try:
                    m = Mesh.load(key[1]) if key[1].endswith(".npz") else load_obj(key[1])
                except OSError as e:
                    raise ProjectError(f"could not load {key[1]}: {e}") from e
            self._keep.append(m)
            self.mesh_ids[key] = len(self.meshes)
            self.meshes.append(m)
        return self.mesh_ids[key]

    def look_at(self, t):
        if not isinstance(t, dict) or t.get("type") != "look_at":
            raise ProjectError("expected a transform.look_at table")
        up = t.get("up")
        return (self.expr(self.need(t, "from")), self.expr(self.need(t, "to")), None if up is None else self.expr(up))


def _w_expr(w: _W, e):
    tag, v = e
    w.u32(tag)
    if tag == 0:
        w.f64(v)
    else:
        w.u32(v)
        w.u32(0)


def _w_opt_expr(w: _W, e):
    if e is None:
        w.u32(0)
    else:
        w.u32(1)
        _w_expr(w, e)


def _w_material(w: _W, m):
    w.u32(m[0])
    _w_opt_expr(w, m[1])


def _w_look_at(w: _W, t):
    _w_expr(w, t[0])
    _w_expr(w, t[1])
    _w_opt_expr(w, t[2])


def _opt_int(t, k):
    v = t.get(k)
    return None if v is None else int(v)


def serialize_project(project: Dict[str, Any], base_dir=None) -> bytes:
    """Decode a project table as `project::Project` (project/mod.rs:103-161) and emit the IR."""
    s = _Serializer(base_dir)
    body = _W()

    # image (project/mod.rs:111-118)
    image = s.need(project, "image")
    body.u32(s.need(image, "width"))
    body.u32(s.need(image, "height"))
    _w_opt_expr(body, None if image.get("filter") is None else s.expr(image["filter"]))
    _w_opt_expr(body, None if image.get("white") is None else s.expr(image["white"]))

    # renderer (project/mod.rs:131-161); `spectrum_bins` is NOT a field
    r = s.need(project, "renderer")
    rtype = {"simple": 0, "bidirectional": 1, "photon_mapping": 2}.get(r.get("type"))
    if rtype is None:
        raise ProjectError(f"unknown renderer type: {r.get('type')!r}")
    body.u32(rtype)
    body.u32(s.need(r, "pixel_samples"))
    for k in ("threads", "bounces", "light_samples", "spectrum_samples", "spectrum_resolution", "tile_size", "light_bounces"):
        body.opt_u32(_opt_int(r, k))

    # camera (project/mod.rs:120-129)
    c = s.need(project, "camera")
    if c.get("type") != "perspective":
        raise ProjectError(f"unknown camera type: {c.get('type')!r}")
    _w_look_at(body, s.look_at(s.need(c, "transform")))
    _w_expr(body, s.expr(s.need(c, "fov")))
    _w_opt_expr(body, None if c.get("focus_distance") is None else s.expr(c["focus_distance"]))
    _w_opt_expr(body, None if c.get("aperture") is None else s.expr(c["aperture"]))

    # world (project/mod.rs:163-203)
    world = s.need(project, "world")
    _w_opt_expr(body, None if world.get("sky") is None else s.expr(world["sky"]))
    objects = list(s.need(world, "objects"))
    body.u32(len(objects))
    for o in objects:
        ty = o.get("type")
        if ty == "sphere":
            body.u32(0)
            _w_expr(body, s.expr(s.need(o, "position")))
            _w_expr(body, s.expr(s.need(o, "radius")))
            _w_opt_expr(body, None if o.get("texture_scale") is None else s.expr(o["texture_scale"]))
            _w_material(body, s.material(s.need(o, "material")))
        elif ty == "plane":
            body.u32(1)
            _w_expr(body, s.expr(s.need(o, "origin")))
            _w_expr(body, s.expr(s.need(o, "normal")))
            _w_opt_expr(body, None if o.get("texture_scale") is None else s.expr(o["texture_scale"]))
            _w_material(body, s.material(s.need(o, "material")))
        elif ty == "ray_marched":
            body.u32(2)
            e = s.need(o, "shape")
            if e.get("type") == "mandelbulb":
                body.u32(0)
                for k in ("iterations", "threshold", "power"):
                    _w_expr(body, s.expr(s.need(e, k)))
                _w_opt_expr(body, None if e.get("constant") is None else s.expr(e["constant"]))
            elif e.get("type") == "quaternion_julia":
                body.u32(1)
                for k in ("iterations", "threshold", "constant", "slice_plane"):
                    _w_expr(body, s.expr(s.need(e, k)))
                name = s.need(s.need(e, "variant"), "name")
                if name not in ("regular", "cubic", "bicomplex"):
                    raise ProjectError(f"unexpected Julia fractal variant: {name}")  # world.rs:162-168
                body.u32({"regular": 0, "cubic": 1, "bicomplex": 2}[name])
            else:
                raise ProjectError(f"unknown estimator type: {e.get('type')!r}")
            b = s.need(o, "bounds")
            if b.get("type") == "box":
                body.u32(0)
                _w_expr(body, s.expr(s.need(b, "min")))
                _w_expr(body, s.expr(s.need(b, "max")))
            elif b.get("type") == "sphere":
                body.u32(1)
                _w_expr(body, s.expr(s.need(b, "position")))
                _w_expr(body, s.expr(s.need(b, "radius")))
            else:
                raise ProjectError(f"unknown bounds type: {b.get('type')!r}")
            _w_material(body, s.material(s.need(o, "material")))
        elif ty == "mesh":
            body.u32(3)
            body.u32(s.mesh(s.need(o, "file")))
            mats = s.need(o, "materials")
            body.u32(len(mats))
            for name, m in mats.items():
                body.string(str(name))
                _w_material(body, s.material(m))
            _w_opt_expr(body, None if o.get("scale") is None else s.expr(o["scale"]))
            if o.get("transform") is None:
                body.u32(0)
            else:
                body.u32(1)
                _w_look_at(body, s.look_at(o["transform"]))
        elif ty == "directional_light":
            body.u32(4)
            for k in ("direction", "width", "color"):
                _w_expr(body, s.expr(s.need(o, k)))
        elif ty == "point_light":
            body.u32(5)
            for k in ("position", "color"):
                _w_expr(body, s.expr(s.need(o, k)))
        else:
            raise ProjectError(f"unknown world object type: {ty!r}")

    # ---- assemble: header, node tables, resources, then the body written above
    w = _W()
    w.u32(IR_MAGIC)
    w.u32(IR_VERSION)

    w.u32(len(s.expr_recs))
    for rec in s.expr_recs:
        w.u32(rec[0])
        if rec[0] == 2:
            w.u32(rec[1])
            for e in rec[2]:
                _w_expr(w, e)
        elif rec[0] in (7, 8, 9):
            w.u32(rec[1])
        else:
            for e in rec[1]:
                _w_expr(w, e)

    w.u32(len(s.mat_recs))
    for rec in s.mat_recs:
        w.u32(rec[0])
        if rec[0] in (0, 1, 2):
            _w_expr(w, rec[1])
        elif rec[0] == 3:
            _w_expr(w, rec[1])
            _w_expr(w, rec[2])
            for e in rec[3]:
                _w_opt_expr(w, e)
        elif rec[0] == 4:
            w.u32(rec[1])
            w.u32(rec[2])
            _w_expr(w, rec[3])
        else:
            w.u32(rec[1])
            w.u32(rec[2])

    w.u32(len(s.spectra))
    for rec in s.spectra:
        if rec[0] == "array":
            w.u32(0)
            w.f32(rec[1])
            w.f32(rec[2])
            w.u32(len(rec[3]))
            w.array(rec[3], np.float32)
        else:
            w.u32(1)
            w.u32(len(rec[1]))
            w.array(rec[1], np.float32)

    w.u32(len(s.color_tex))
    for t in s.color_tex:
        w.u32(t.shape[1])
        w.u32(t.shape[0])
        w.array(t, np.float32)
    w.u32(len(s.mono_tex))
    for t in s.mono_tex:
        w.u32(t.shape[1])
        w.u32(t.shape[0])
        w.array(t, np.float32)

    w.u32(len(s.meshes))
    for m in s.meshes:
        w.u32(len(m.position))
        w.array(m.position, np.float32)
        w.u32(len(m.texture))
        w.array(m.texture, np.float32)
        w.u32(len(m.normal))
        w.array(m.normal, np.float32)
        w.u32(len(m.objects))
        for name, tris in m.objects:
            w.string(name)
            w.u32(len(tris))
            w.array(tris, np.int32)

    # constant resources (the reference compiles these in: build.rs; rgb.rs, xyz.rs, light_source.rs)
    t = s.tables
    w.f32(t["burns_min"])
    w.f32(t["burns_max"])
    w.u32(len(t["burns_rgb"]))
    w.array(t["burns_rgb"], np.float32)
    w.f32(t["xyz_min"])
    w.f32(t["xyz_max"])
    w.u32(len(t["xyz"]))
    w.array(t["xyz"], np.float32)
    w.f32(t["illum_min"])
    w.f32(t["illum_max"])
    w.u32(len(t["d65"]))
    w.array(t["d65"], np.float32)

    return w.bytes() + body.bytes()


def project_summary(project) -> str:
    r = project["renderer"]
    i = project["image"]
    return f"{i['width']}x{i['height']} {r['type']} spp={r['pixel_samples']}"


__all__ = [
    "Node", "Expr", "mix", "fresnel", "vector", "blackbody", "spectrum", "rgb", "texture", "shape", "ray_marched",
    "quaternion_julia", "bounds", "material", "light_source", "transform", "camera", "renderer", "light", "Mesh",
    "Image", "parse_obj", "load_obj", "serialize_project", "ProjectError", "builtin_tables", "srgb_to_linear",
    "IR_MAGIC", "IR_VERSION",
]
_ = math  # keep import for users of this namespace
