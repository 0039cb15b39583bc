"""ctypes wrapper around oracle/liboracle.so - the CPU restatement of pyrite's render path.

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs only.  The product package (pyrite_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent.parent / "oracle"
LIB_PATH = ORACLE_DIR / "liboracle.so"

RAY_DTYPE = np.dtype([("o", np.float32, 3), ("pad0", np.float32), ("d", np.float32, 3), ("pad1", np.float32)])
HIT_DTYPE = np.dtype([("prim_id", np.uint32), ("kind", np.uint32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])

KIND_MISS, KIND_PLANE, KIND_TRIANGLE, KIND_SPHERE, KIND_RAY_MARCHED = 0, 1, 2, 3, 4


class Info(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "width", "height", "bins", "algorithm", "pixel_samples", "bounces", "light_samples", "spectrum_samples", "light_bounces",
        "tile_size", "n_objects", "n_planes", "n_lights", "n_bvh_nodes", "n_materials", "threads")]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "nodes", "leaves", "path_samples", "de_evals", "de_iters")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class RenderOpts(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("rng_mode", C.c_int32), ("eager_emissive_draw", C.c_int32), ("spp_override", C.c_uint32),
                ("sample_offset", C.c_uint32), ("sample_stride", C.c_uint32), ("threads", C.c_int32), ("cas_attempts", C.c_int32),
                ("reset_film", C.c_int32)]


# variants: "glibc" = the reference's arithmetic (parity tests); "double" = shading-side libm in double, rounded once, like
# the device code (attribution of film differences); "native" = -O3 -march=native, built on the box that runs it
# (bench.py's CPU baseline legs only: the reference's own release profile)
VARIANTS = {"glibc": "liboracle.so", "double": "liboracle_dbl.so", "native": "liboracle_native.so"}


def build(force: bool = False, variant: str = "glibc") -> Path:
    target = ORACLE_DIR / VARIANTS[variant]
    srcs = list(ORACLE_DIR.glob("*.hpp")) + list(ORACLE_DIR.glob("*.cpp")) + [ORACLE_DIR / "Makefile"]
    stale = not target.exists() or any(s.stat().st_mtime > target.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(ORACLE_DIR), target.name] + (["-B"] if force else ["-s"]), check=True, capture_output=True)
    return target


_libs = {}


def lib(variant: str = "glibc"):
    if variant not in _libs:
        L = C.CDLL(str(build(variant=variant)))
        L.pyro_libm_mode.restype = C.c_int
        assert L.pyro_libm_mode() == (1 if variant == "double" else 0)
        L.pyro_last_error.restype = C.c_char_p
        L.pyro_last_render_seconds.restype = C.c_double
        L.pyro_last_render_seconds.argtypes = [C.c_void_p]
        L.pyro_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.pyro_free.argtypes = [C.c_void_p]
        L.pyro_info.argtypes = [C.c_void_p, C.POINTER(Info)]
        L.pyro_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.POINTER(Counters)]
        L.pyro_gen_rays.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_uint64, C.c_void_p]
        L.pyro_render.argtypes = [C.c_void_p, C.POINTER(RenderOpts)]
        L.pyro_counters.argtypes = [C.c_void_p, C.POINTER(Counters), C.c_int]
        L.pyro_film_download.argtypes = [C.c_void_p, C.c_void_p]
        L.pyro_film_upload.argtypes = [C.c_void_p, C.c_void_p]
        L.pyro_film_expose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.pyro_film_develop.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
        L.pyro_bvh_leaf_order.argtypes = [C.c_void_p, C.c_void_p]
        L.pyro_camera_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pyro_render_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]
        L.pyro_debug_path.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _libs[variant] = L
    return _libs[variant]


class OracleError(RuntimeError):
    pass


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One loaded project (the equivalent of the reference after `parse_project`, main.rs:111-134)."""

    def __init__(self, ir: bytes, variant: str = "glibc"):
        self.L = lib(variant)
        self.h = C.c_void_p()
        buf = (C.c_char * len(ir)).from_buffer_copy(ir)
        if self.L.pyro_load(buf, len(ir), C.byref(self.h)) != 0:
            raise OracleError(self.L.pyro_last_error().decode())
        self.info = Info()
        self.L.pyro_info(self.h, C.byref(self.info))
        self.threads = os.cpu_count() or 1

    def close(self):
        if self.h:
            self.L.pyro_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise OracleError(self.L.pyro_last_error().decode())

    def trace(self, rays: np.ndarray, threads: int | None = None):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(len(rays), dtype=HIT_DTYPE)
        c = Counters()
        self._check(self.L.pyro_trace(self.h, _ptr(rays), len(rays), _ptr(hits), threads or self.threads, C.byref(c)))
        return hits, c.as_dict()

    def gen_rays(self, kind: int, n: int, seed: int = 0) -> np.ndarray:
        rays = np.zeros(n, dtype=RAY_DTYPE)
        self._check(self.L.pyro_gen_rays(self.h, kind, n, seed, _ptr(rays)))
        return rays

    def render(self, seed=1, rng_mode=1, eager_emissive_draw=True, spp=0, sample_offset=0, sample_stride=1, threads=None,
               cas_attempts=0, reset_film=True) -> float:
        o = RenderOpts(seed, rng_mode, int(eager_emissive_draw), spp, sample_offset, sample_stride, threads or self.threads,
                       cas_attempts, int(reset_film))
        self._check(self.L.pyro_render(self.h, C.byref(o)))
        return float(self.L.pyro_last_render_seconds(self.h))

    def render_sample(self, seed: int, tile: int, sample: int, reset_film: bool = True):
        """Exactly one path sample (tile, sample) of the project's integrator into the film."""
        self._check(self.L.pyro_render_sample(self.h, seed, tile, sample, int(reset_film)))

    def counters(self, reset=False) -> dict:
        c = Counters()
        self.L.pyro_counters(self.h, C.byref(c), int(reset))
        return c.as_dict()

    def film(self) -> np.ndarray:
        i = self.info
        out = np.empty((i.height, i.width, i.bins, 2), dtype=np.float32)
        self._check(self.L.pyro_film_download(self.h, _ptr(out)))
        return out

    def set_film(self, film: np.ndarray):
        film = np.ascontiguousarray(film, dtype=np.float32)
        self._check(self.L.pyro_film_upload(self.h, _ptr(film)))

    def expose(self, positions: np.ndarray, samples: np.ndarray):
        positions = np.ascontiguousarray(positions, dtype=np.float32)
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        self._check(self.L.pyro_film_expose(self.h, _ptr(positions), _ptr(samples), len(positions)))

    def develop(self, step_size: float = 2.0, threads=None):
        i = self.info
        xyz = np.empty((i.height, i.width, 3), dtype=np.float32)
        srgb = np.empty((i.height, i.width, 3), dtype=np.uint8)
        self._check(self.L.pyro_film_develop(self.h, step_size, _ptr(xyz), _ptr(srgb), threads or self.threads))
        return xyz, srgb

    def debug_path(self, seed, tile, sample, max_bounces=64, eager_emissive_draw=True):
        """One `render_tile` iteration of simple.rs with per-bounce records: (records[n, 20] u32, exposed[m, 2] f32, position[2])."""
        rec = np.zeros((max_bounces, 20), np.uint32)
        exposed = np.zeros((16, 2), np.float32)
        pos = np.zeros(2, np.float32)
        nb, ne = C.c_uint32(), C.c_uint32()
        self._check(self.L.pyro_debug_path(self.h, seed, int(eager_emissive_draw), tile, sample, max_bounces, _ptr(rec), C.byref(nb), _ptr(exposed), C.byref(ne), _ptr(pos)))
        return rec[:min(nb.value, max_bounces)], exposed[:ne.value], pos

    def bvh_leaf_order(self) -> np.ndarray:
        out = np.empty(self.info.n_objects, dtype=np.uint32)
        self.L.pyro_bvh_leaf_order(self.h, _ptr(out))
        return out

    def camera_sample(self, seed: int, tile: int, sample: int):
        pos = np.zeros(2, np.float32)
        ray = np.zeros(1, RAY_DTYPE)
        wl = np.zeros(self.info.spectrum_samples, np.float32)
        hero = C.c_uint32()
        self._check(self.L.pyro_camera_sample(self.h, seed, tile, sample, _ptr(pos), _ptr(ray), _ptr(wl), C.byref(hero)))
        return pos, ray[0], wl, int(hero.value)
