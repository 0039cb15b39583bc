"""ctypes wrapper around tests/_build/libhostemu.so (tests/host_emu.cpp): the product's scene builder
and __host__ __device__ stage functions run on the CPU.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from oracle_lib import HIT_DTYPE, RAY_DTYPE

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "tests" / "_build"
LIB = BUILD / "libhostemu.so"
SOURCES = [ROOT / "tests" / "host_emu.cpp", ROOT / "pyrite_b200" / "csrc" / "scene_build.cpp"]
HEADERS = list((ROOT / "pyrite_b200" / "csrc").glob("*.h*")) + list((ROOT / "pyrite_b200" / "csrc").glob("*.cuh"))


def build(force=False):
    BUILD.mkdir(exist_ok=True)
    deps = SOURCES + HEADERS
    if force or not LIB.exists() or any(d.stat().st_mtime > LIB.stat().st_mtime for d in deps):
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", str(LIB)]
        subprocess.run(cmd + [str(s) for s in SOURCES], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.emu_last_error.restype = C.c_char_p
        L.emu_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.emu_load_with.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        L.emu_bvh_digest.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_free.argtypes = [C.c_void_p]
        L.emu_info.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_leaf_order.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.emu_render.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.emu_rays.argtypes = [C.c_void_p]
        L.emu_rays.restype = C.c_uint64
        L.emu_film.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_set_film.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_expose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.emu_develop.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        L.emu_run_program.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.emu_camera_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_debug_path.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class EmuError(RuntimeError):
    pass


class Emu:
    def __init__(self, ir: bytes, shape, level_sync_bvh: bool = False):
        """shape = (height, width, bins) of the film (from the oracle's / product's project info).  `level_sync_bvh`: build the BVH
        with the level-synchronous algorithm of the GPU builder (bvh_build_core.hpp) instead of the depth-first host builder."""
        self.L = lib()
        self.h = C.c_void_p()
        buf = (C.c_char * len(ir)).from_buffer_copy(ir)
        if self.L.emu_load_with(buf, len(ir), 1 if level_sync_bvh else 0, C.byref(self.h)) != 0:
            raise EmuError(self.L.emu_last_error().decode())
        self.shape = tuple(shape)
        info = np.zeros(9, np.uint32)
        self.L.emu_info(self.h, _ptr(info))
        self.info = dict(zip(("n_objects", "n_planes", "n_lamps", "n_nodes", "n_materials", "n_programs", "n_instr", "n_tiles", "bvh_depth"), map(int, info)))

    def bvh_digest(self):
        out = np.zeros(2, np.uint64)
        self.L.emu_bvh_digest(self.h, _ptr(out))
        return int(out[0]), int(out[1])

    def __del__(self):
        try:
            if self.h:
                self.L.emu_free(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def leaf_order(self):
        out = np.zeros(self.info["n_objects"], np.uint32)
        self.L.emu_leaf_order(self.h, _ptr(out))
        return out

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(len(rays), dtype=HIT_DTYPE)
        st = np.zeros(3, np.uint64)
        self.L.emu_trace(self.h, _ptr(rays), len(rays), _ptr(hits), _ptr(st))
        return hits, {"rays": int(st[0]), "nodes": int(st[1]), "leaves": int(st[2])}

    def render(self, seed=1, spp=0, sample_offset=0, sample_stride=1, reset_film=True):
        before = int(self.L.emu_rays(self.h))
        if self.L.emu_render(self.h, seed, spp, sample_offset, sample_stride, int(reset_film)) != 0:
            raise EmuError(self.L.emu_last_error().decode())
        self.rays_last = int(self.L.emu_rays(self.h)) - before
        return self.rays_last

    def film(self):
        out = np.empty(self.shape + (2,), np.float32)
        self.L.emu_film(self.h, _ptr(out))
        return out

    def set_film(self, film):
        film = np.ascontiguousarray(film, np.float32)
        self.L.emu_set_film(self.h, _ptr(film))

    def expose(self, positions, samples):
        positions = np.ascontiguousarray(positions, np.float32)
        samples = np.ascontiguousarray(samples, np.float32)
        self.L.emu_expose(self.h, _ptr(positions), _ptr(samples), len(positions))

    def develop(self, step_size=2.0):
        h, w, _ = self.shape
        xyz = np.empty((h, w, 3), np.float32)
        srgb = np.empty((h, w, 3), np.uint8)
        self.L.emu_develop(self.h, step_size, _ptr(xyz), _ptr(srgb))
        return xyz, srgb

    def debug_path(self, seed, tile, sample, max_bounces=64):
        """One path sample of the camera-to-light integrator, depth-first: (records[n, 20] u32, exposed[m, 2] f32, position[2])."""
        rec = np.zeros((max_bounces, 20), np.uint32)
        exposed = np.zeros((16, 2), np.float32)
        pos = np.zeros(2, np.float32)
        nb, ne = C.c_uint32(), C.c_uint32()
        self.L.emu_debug_path(self.h, seed, tile, sample, max_bounces, _ptr(rec), C.byref(nb), _ptr(exposed), C.byref(ne), _ptr(pos))
        return rec[:min(nb.value, max_bounces)], exposed[:ne.value], pos

    def camera_sample(self, seed, tile, sample, spectrum_samples):
        pos = np.zeros(2, np.float32)
        ray = np.zeros(1, RAY_DTYPE)
        wl = np.zeros(16, np.float32)
        hero = C.c_uint32()
        self.L.emu_camera_sample(self.h, seed, tile, sample, _ptr(pos), _ptr(ray), _ptr(wl), C.byref(hero))
        return pos, ray[0], wl[:spectrum_samples], int(hero.value)
