"""Host-side mirror of the reference's render interface over the C ABI (include/pyrite_b200.h).

The reference's `main` (pyrite/src/main.rs:52-109) does `load_project` -> `parse_project` ->
`render` -> develop.  `Renderer` below exposes the same steps on one CUDA device:

    r = Renderer(device=0)                       # pyr_init
    r.load(project_table)                        # load_project + parse_project   (pyr_project_load)
    r.render(seed=1)                             # Renderer::render               (pyr_render)
    xyz, srgb = r.develop()                      # main.rs:313-327                (pyr_film_develop)
    hits = r.trace(rays)                         # World::intersect on a batch    (pyr_trace)

Everything that computes runs in libpyrite_b200.so on the GPU.  There is no CPU fallback: if the
library is missing or no CUDA device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import numpy as np

from .project import serialize_project

LIB_PATH = Path(__file__).resolve().parent / "libpyrite_b200.so"

RAY_DTYPE = np.dtype([("o", np.float32, 3), ("pad0", np.float32), ("d", np.float32, 3), ("pad1", np.float32)])
HIT_DTYPE = np.dtype([("prim_id", np.uint32), ("kind", np.uint32), ("t", np.float32), ("u", np.float32), ("v", np.float32)])
KIND_MISS, KIND_PLANE, KIND_TRIANGLE, KIND_SPHERE, KIND_RAY_MARCHED = 0, 1, 2, 3, 4
RENDER_STATS = 1
RENDER_TIMING = 2

EXPORTS = [
    "pyr_init", "pyr_shutdown", "pyr_stream_set", "pyr_last_error", "pyr_project_load", "pyr_project_info_get", "pyr_trace", "pyr_trace_device",
    "pyr_trace_stats", "pyr_bvh_leaf_order", "pyr_bvh_digest", "pyr_render", "pyr_film_expose", "pyr_film_clear", "pyr_film_download", "pyr_film_upload",
    "pyr_film_device_ptr", "pyr_comm_unique_id", "pyr_comm_init", "pyr_comm_init_async", "pyr_film_reduce", "pyr_comm_destroy", "pyr_film_develop", "pyr_camera_sample", "pyr_debug_path", "pyr_counters_get", "pyr_version",
]


class PyriteError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"pyrite_b200 error {status}: {message}")
        self.status = status


class ProjectInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "width", "height", "bins", "algorithm", "pixel_samples", "bounces", "light_samples", "spectrum_samples", "light_bounces",
        "tile_size", "n_objects", "n_planes", "n_lights", "n_bvh_nodes", "n_materials", "n_ray_marched")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class RenderParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("spp_override", C.c_uint32), ("sample_offset", C.c_uint32), ("sample_stride", C.c_uint32),
                ("reset_film", C.c_uint32), ("pool_paths", C.c_uint32), ("flags", C.c_uint32), ("tile_filter", C.c_uint32), ("reserved", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "path_samples", "nodes_visited", "leaves_tested", "de_evals", "de_iterations",
                                          "wavefront_iterations", "kernel_launches")] + [
        ("render_seconds", C.c_double), ("trace_seconds", C.c_double), ("shade_seconds", C.c_double),
        ("trace_launches", C.c_uint64), ("shade_launches", C.c_uint64), ("node_fetches", C.c_uint64), ("path_rays", C.c_uint64),
        ("march_iterations", C.c_uint64), ("julia_iterations", C.c_uint64)]

    def as_dict(self):
        return {n: (float if n.endswith("_seconds") else int)(getattr(self, n)) for n, _ in self._fields_}


PROGRESS_CB = C.CFUNCTYPE(C.c_int, C.c_uint8, C.c_char_p, C.c_void_p)

_lib = None


def load_library(path: Optional[Path] = None):
    """dlopen the product library and declare its signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise ImportError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(make -C pyrite_b200/csrc).  There is no CPU fallback.")
    L = C.CDLL(str(p))
    vp, sz = C.c_void_p, C.c_size_t
    L.pyr_init.argtypes = [C.c_int32, C.POINTER(vp)]
    L.pyr_shutdown.argtypes = [vp]
    L.pyr_shutdown.restype = None
    L.pyr_stream_set.argtypes = [vp, vp]
    L.pyr_last_error.argtypes = [vp]
    L.pyr_last_error.restype = C.c_char_p
    L.pyr_version.restype = C.c_char_p
    L.pyr_project_load.argtypes = [vp, vp, sz]
    L.pyr_project_info_get.argtypes = [vp, C.POINTER(ProjectInfo)]
    L.pyr_trace.argtypes = [vp, vp, sz, vp]
    L.pyr_trace_stats.argtypes = [vp, vp, sz, vp]
    L.pyr_trace_device.argtypes = [vp, vp, sz, vp, C.c_uint32]
    L.pyr_bvh_leaf_order.argtypes = [vp, vp]
    L.pyr_bvh_digest.argtypes = [vp, vp]
    L.pyr_render.argtypes = [vp, C.POINTER(RenderParams), PROGRESS_CB, vp]
    L.pyr_film_expose.argtypes = [vp, vp, vp, sz]
    L.pyr_film_clear.argtypes = [vp]
    L.pyr_film_download.argtypes = [vp, vp]
    L.pyr_film_upload.argtypes = [vp, vp]
    L.pyr_film_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.pyr_film_develop.argtypes = [vp, C.c_float, vp, vp]
    L.pyr_comm_unique_id.argtypes = [vp]
    L.pyr_comm_init.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.pyr_comm_init_async.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.pyr_film_reduce.argtypes = [vp, C.c_int32]
    L.pyr_comm_destroy.argtypes = [vp]
    L.pyr_camera_sample.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint64, vp, vp, vp, vp]
    L.pyr_debug_path.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, vp, vp, vp, vp, vp]
    L.pyr_counters_get.argtypes = [vp, C.POINTER(Counters), C.c_int32]
    for name in EXPORTS:
        if name not in ("pyr_shutdown", "pyr_last_error", "pyr_version"):
            getattr(L, name).restype = C.c_int32
    if path is None:
        _lib = L
    return L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class DeviceFilm:
    """The context's film in device memory, exposed through `__cuda_array_interface__` so that
    `torch.as_tensor(film, device="cuda")` aliases it without a copy (the NCCL reduction of the
    multi-GPU path runs on that alias)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2, "strides": None}


class Renderer:
    def __init__(self, device: int = 0):
        self.L = load_library()
        self.h = C.c_void_p()
        status = self.L.pyr_init(device, C.byref(self.h))
        if status != 0:
            raise PyriteError(status, self.L.pyr_last_error(None).decode())
        self.device = device
        self.info: Optional[ProjectInfo] = None

    # -- lifetime
    def close(self):
        if getattr(self, "h", None):
            self.L.pyr_shutdown(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, status: int):
        if status != 0:
            raise PyriteError(status, self.L.pyr_last_error(self.h).decode())

    def set_stream(self, cuda_stream: Optional[int]):
        """Launch on the caller's CUDA stream (e.g. `torch.cuda.current_stream().cuda_stream`); None = own stream."""
        self._check(self.L.pyr_stream_set(self.h, C.c_void_p(cuda_stream or 0)))

    # -- project
    def load(self, project, base_dir=None):
        """`project` is a project table (pyrite_b200.project DSL) or an already serialised IR blob."""
        ir = project if isinstance(project, (bytes, bytearray)) else serialize_project(project, base_dir)
        buf = (C.c_char * len(ir)).from_buffer_copy(ir)
        self._check(self.L.pyr_project_load(self.h, buf, len(ir)))
        self.info = ProjectInfo()
        self._check(self.L.pyr_project_info_get(self.h, C.byref(self.info)))
        return self.info

    @property
    def film_shape(self):
        i = self.info
        return (i.height, i.width, i.bins, 2)

    # -- World::intersect
    def trace(self, rays: np.ndarray, stats: bool = False) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(len(rays), dtype=HIT_DTYPE)
        fn = self.L.pyr_trace_stats if stats else self.L.pyr_trace
        self._check(fn(self.h, _ptr(rays), len(rays), _ptr(hits)))
        return hits

    def trace_device(self, d_rays: int, n: int, d_hits: int, repeat: int = 1) -> float:
        """Trace `n` rays resident at device address `d_rays`; returns the device seconds of all repeats."""
        self._check(self.L.pyr_trace_device(self.h, C.c_void_p(d_rays), n, C.c_void_p(d_hits), repeat))
        return self.counters()["render_seconds"]

    def bvh_digest(self) -> dict:
        """Test hook: whether the BVH was built on the GPU, digests of the 4-wide nodes and of the leaf pre-order, and the
        seconds the last `load` took inside the library."""
        out = (C.c_uint64 * 4)()
        self._check(self.L.pyr_bvh_digest(self.h, out))
        return {"built_on_gpu": bool(out[0]), "nodes": int(out[1]), "leaf_order": int(out[2]), "load_seconds": out[3] * 1e-6}

    def bvh_leaf_order(self) -> np.ndarray:
        out = np.empty(self.info.n_objects, dtype=np.uint32)
        self._check(self.L.pyr_bvh_leaf_order(self.h, _ptr(out)))
        return out

    # -- Renderer::render
    def render(self, seed: int = 0, spp: int = 0, sample_offset: int = 0, sample_stride: int = 1, reset_film: bool = True,
               pool_paths: int = 0, stats: bool = False, timing: bool = False, progress=None, only_tile=None) -> float:
        """Runs the wavefront pipeline into the film; returns the device seconds it took.  `only_tile`: render one tile only (diagnostics)."""
        p = RenderParams(seed, spp, sample_offset, sample_stride, int(reset_film), pool_paths, (RENDER_STATS if stats else 0) | (RENDER_TIMING if timing else 0),
                         0 if only_tile is None else only_tile + 1, 0)
        if progress is None:
            cb = PROGRESS_CB()
        else:
            cb = PROGRESS_CB(lambda pct, msg, _user: int(bool(progress(int(pct), msg.decode() if msg else ""))))
        self._check(self.L.pyr_render(self.h, C.byref(p), cb, None))
        return self.counters()["render_seconds"]

    # -- film
    def expose(self, positions: np.ndarray, samples: np.ndarray):
        positions = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 2)
        samples = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 3)
        self._check(self.L.pyr_film_expose(self.h, _ptr(positions), _ptr(samples), len(positions)))

    def clear_film(self):
        self._check(self.L.pyr_film_clear(self.h))

    def film(self) -> np.ndarray:
        out = np.empty(self.film_shape, dtype=np.float32)
        self._check(self.L.pyr_film_download(self.h, _ptr(out)))
        return out

    def set_film(self, film: np.ndarray):
        film = np.ascontiguousarray(film, dtype=np.float32)
        assert film.shape == self.film_shape
        self._check(self.L.pyr_film_upload(self.h, _ptr(film)))

    def film_device(self) -> DeviceFilm:
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        self._check(self.L.pyr_film_device_ptr(self.h, C.byref(ptr), C.byref(nbytes)))
        assert nbytes.value == int(np.prod(self.film_shape)) * 4
        return DeviceFilm(ptr.value, self.film_shape)

    # -- multi-GPU: the library-owned NCCL communicator and the one film reduction
    @staticmethod
    def comm_unique_id() -> bytes:
        """Rank 0: a fresh communicator id (128 bytes) to hand to every rank's `comm_init`."""
        L = load_library()
        buf = (C.c_uint8 * 128)()
        status = L.pyr_comm_unique_id(buf)
        if status != 0:
            raise PyriteError(status, L.pyr_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, n_ranks: int, rank: int, comm_id: bytes, wait: bool = True):
        """Every rank, collectively.  `wait=False` (`pyr_comm_init_async`) returns at once and lets the library set the
        communicator up while this rank renders; `film_reduce` waits for it."""
        assert len(comm_id) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(comm_id)
        self._check((self.L.pyr_comm_init if wait else self.L.pyr_comm_init_async)(self.h, n_ranks, rank, buf))

    def film_reduce(self, root: int = 0):
        """Collective: film := sum of the ranks' films on `root` (root < 0: everywhere)."""
        self._check(self.L.pyr_film_reduce(self.h, root))

    def comm_destroy(self):
        self._check(self.L.pyr_comm_destroy(self.h))

    def develop(self, step_size: float = 2.0, want_xyz: bool = True, want_srgb: bool = True, out_xyz=None, out_srgb=None):
        """Film -> (XYZ f32, sRGB u8) images on the host.  `out_xyz` / `out_srgb`: caller-owned C-contiguous arrays of the
        image shape to download into (page-locked ones make the copies DMA transfers); otherwise fresh arrays."""
        i = self.info
        xyz = srgb = None
        if want_xyz:
            xyz = out_xyz if out_xyz is not None else np.empty((i.height, i.width, 3), dtype=np.float32)
            if xyz.dtype != np.float32 or xyz.shape != (i.height, i.width, 3) or not xyz.flags.c_contiguous:
                raise ValueError("out_xyz must be a C-contiguous float32 array of shape (height, width, 3)")
        if want_srgb:
            srgb = out_srgb if out_srgb is not None else np.empty((i.height, i.width, 3), dtype=np.uint8)
            if srgb.dtype != np.uint8 or srgb.shape != (i.height, i.width, 3) or not srgb.flags.c_contiguous:
                raise ValueError("out_srgb must be a C-contiguous uint8 array of shape (height, width, 3)")
        self._check(self.L.pyr_film_develop(self.h, step_size, _ptr(xyz) if want_xyz else None, _ptr(srgb) if want_srgb else None))
        return xyz, srgb

    # -- camera seam
    def camera_sample(self, seed: int, tile: int, sample: int):
        pos = np.zeros(2, np.float32)
        ray = np.zeros(1, RAY_DTYPE)
        wl = np.zeros(self.info.spectrum_samples, np.float32)
        hero = C.c_uint32()
        self._check(self.L.pyr_camera_sample(self.h, seed, tile, sample, _ptr(pos), _ptr(ray), _ptr(wl), C.byref(hero)))
        return pos, ray[0], wl, int(hero.value)

    def debug_path(self, seed: int, tile: int, sample: int, max_bounces: int = 64):
        """Diagnostic: one path sample of the camera-to-light integrator run depth-first by one GPU thread.
        Returns (records[n, 20] uint32, exposed[m, 2] float32 (brightness, wavelength), position[2])."""
        rec = np.zeros((max_bounces, 20), np.uint32)
        exposed = np.zeros((16, 2), np.float32)
        pos = np.zeros(2, np.float32)
        nb, ne = C.c_uint32(), C.c_uint32()
        self._check(self.L.pyr_debug_path(self.h, seed, tile, sample, max_bounces, _ptr(rec), C.byref(nb), _ptr(exposed), C.byref(ne), _ptr(pos)))
        return rec[:min(nb.value, max_bounces)], exposed[:ne.value], pos

    # -- counters
    def counters(self, reset: bool = False) -> dict:
        c = Counters()
        self._check(self.L.pyr_counters_get(self.h, C.byref(c), int(reset)))
        return c.as_dict()

    def version(self) -> str:
        return self.L.pyr_version().decode()
