#!/usr/bin/env python3
"""Benchmark of the render hot path (BASELINE.json metric: Mrays/s and spectral path samples/s).

Default workload: BASELINE config C2 - the ~870k-triangle dragon stand-in, 1920x1080, camera-to-light integrator with
S=10 wavelengths, B=8 bounces, L=4 light samples.  One *step* = `--spp-per-step` sample passes over the whole image per
GPU (256 / spp-per-step steps make the full 256-spp config), so throughput does not depend on how many steps are
timed; at N > 1 every rank renders its own passes (weak scaling) and the films are summed once (NCCL, inside the library).
A ray = one World::intersect call, a path sample = one render_tile iteration.

    python bench.py --gpus 1 --steps 8 --warmup 3            # product arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps 3 --warmup 1    # CPU arm: the oracle port on the host cores
    python bench.py --config C5 --spp 64                     # a FIXED job (strong scaling): load -> render -> reduce -> develop

Every run also measures a fixed C5 job (4K bidirectional, `--strong-spp` = 256 sample passes, a quarter of the 1024-spp target, split
over the ranks) and
reports it under "strong_scaling", so that the 1/2/4/8-GPU sweep of the default command carries strong-scaling numbers
next to the weak-scaling headline.  Prints ONE JSON line (rank 0).  DESIGN.md §8 defines every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

NODE_BYTES, PRIM_BYTES, RAY_BYTES, HIT_BYTES = 128, 48, 32, 32  # Node4, Prim, Ray, Hit records (device_types.h)
# bytes per ray of the reference-order walk = 32 (ray) + 20 (hit) + Vn * 32 (binary node) + Vt * 48 (triangle)
REF_RAY_IO, REF_NODE_BYTES, REF_TRI_BYTES = 52, 32, 48

CONFIGS = {
    "C1": dict(workload="C1 Cornell box (box.obj, 36 tris), 512x512, simple integrator S=10 B=4 L=1", spp=64),
    "C2": dict(workload="C2 dragon stand-in 871,200 tris + 3 planes + point light, 1920x1080, simple integrator S=10 B=8 L=4", spp=256),
    "C3": dict(workload="C3 diamonds.lua (dispersive refraction, thin lens, S=1 B=256) at 1920x1080", spp=200),
    "C4": dict(workload="C4 Mandelbulb + cubic quaternion Julia (sphere tracing), 3840x2160, simple integrator S=10 B=8 L=4", spp=512),
    "C5": dict(workload="C5 textured Cornell box + dragon stand-in 871,236 tris, 3840x2160, bidirectional S=10 B=6 LB=6 L=1", spp=1024),
}


def make_scene(config: str, small: bool):
    from pyrite_b200 import scenes

    if config == "C1":
        return scenes.cornell(width=512, height=512, spp=64)
    if config == "C2":
        if small:
            return scenes.dragon(width=480, height=270, mesh=scenes.dragon_mesh(400, 50))
        return scenes.dragon()
    if config == "C3":
        return scenes.diamonds(width=480 if small else 1920, height=270 if small else 1080)
    if config == "C4":
        return scenes.fractals(width=480 if small else 3840, height=270 if small else 2160)
    if config == "C5":
        if small:
            return scenes.bdpt_cornell_dragon(width=480, height=270, mesh=scenes.dragon_mesh(400, 50))
        return scenes.bdpt_cornell_dragon()
    raise SystemExit(f"unknown config {config}")


def build_project(config: str, small: bool) -> bytes:
    from pyrite_b200 import project

    return project.serialize_project(make_scene(config, small))


def workload_name(config: str, small: bool) -> str:
    return CONFIGS[config]["workload"] + (" [--small: 480x270, 40k-triangle mesh]" if small else "")


def config_dict(args, world: int) -> dict:
    """The same dict in both arms (the driver compares them)."""
    fixed = args.config is not None
    return {
        "workload": workload_name(args.config or "C2", args.small),
        "step": (f"fixed job: {args.spp or CONFIGS[args.config]['spp']} spp in total" if fixed else
                 f"{args.spp_per_step} spp per GPU per step ({256 // max(args.spp_per_step, 1)} steps = the 256-spp config on 1 GPU); the CPU arm renders a bounded sample of the same passes"),
        "parallelism": "sample-pass sharding, one process per GPU, one NCCL film reduce (pyr_film_reduce) at the end",
        "l2": "working set (film >= 1.06 GB + path pool + 150 MB BVH) exceeds the 126 MB L2; no flush needed",
    }


class ClockSampler:
    """Samples nvidia-smi during the timed region (the recipe's clocks line)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    """HBM GB/s from the driver-written MEASURED_PEAKS.json; FP32 FFMA TFLOP/s from tools/ffma_peak (profiles/ffma_peak.json)."""
    hbm, hbm_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            hbm, hbm_src = float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    ffma, ffma_src = 74.45, "nominal (148 SMs x 128 lanes x 2 x 1.965 GHz)"
    exe = ROOT / "tools" / "build" / "ffma_peak"
    try:
        if exe.exists():
            ffma, ffma_src = float(json.loads(subprocess.run([str(exe)], capture_output=True, text=True, timeout=60).stdout)["ffma_tflops"]), "measured now (tools/ffma_peak.cu)"
        else:
            ffma, ffma_src = float(json.loads((ROOT / "profiles" / "ffma_peak.json").read_text())["ffma_tflops"]), "measured (profiles/ffma_peak.json, tools/ffma_peak.cu)"
    except Exception:
        pass
    return hbm, hbm_src, ffma, ffma_src


def oracle_variant():
    """The CPU legs use the oracle built like the reference's release profile (opt-level 3, LTO, target-cpu=native), compiled
    on the box that runs it; the portable -O2 build is what the parity tests use."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib

    try:
        oracle_lib.build(force=True, variant="native")   # always rebuilt here: -march=native code must never travel between hosts
        return "native", "g++ -O3 -march=native -ffp-contract=off (one translation unit), built on this host"
    except Exception as e:  # no compiler on the box: fall back to the shipped portable build and say so
        return "glibc", f"portable -O2 build (native build failed: {type(e).__name__})"


def oracle_leg(ir: bytes, threads: int, spp: int, fraction_den: int = 1, seed: int = 1, variant: str = "glibc"):
    """The CPU restatement of pyrite's renderer (oracle/) on `spp`/fraction_den sample passes -> (seconds, counters)."""
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_lib import Oracle

    o = Oracle(ir, variant)
    o.counters(reset=True)
    secs = o.render(seed=seed, spp=spp, sample_offset=0, sample_stride=fraction_den, threads=threads, cas_attempts=5)
    c = o.counters()
    o.close()
    return secs, c


def run_reference(args):
    """--impl reference: pyrite's own CPU algorithm (the oracle port; the Rust crate cannot be built in this image) with all
    host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_lib import Oracle

    threads = os.cpu_count() or 1
    config = args.config or "C2"
    ir = build_project(config, args.small)
    variant, how = oracle_variant()
    o = Oracle(ir, variant)
    den = args.cpu_fraction
    for w in range(args.warmup):
        o.render(seed=100 + w, spp=1, sample_offset=w % den, sample_stride=den * 4, threads=threads, cas_attempts=5)
    o.counters(reset=True)
    total = 0.0
    for k in range(args.steps):
        total += o.render(seed=k, spp=1, sample_offset=k % den, sample_stride=den, threads=threads, cas_attempts=5, reset_film=(k == 0))
    c = o.counters()
    mrays = c["rays"] / total / 1e6
    sample = f"{args.steps} steps x 1/{den} of one sample pass over the whole image ({c['path_samples']} path samples, {c['rays']} rays, {total:.1f} s)"
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "path_samples_per_s": c["path_samples"] / total,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong" if args.config else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, 1),
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample, "build": how,
                         "reference_order_nodes_per_ray": c["nodes"] / max(c["rays"], 1), "reference_order_leaves_per_ray": c["leaves"] / max(c["rays"], 1)},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host": "CPU only",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def dist_setup():
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def barrier(world):
    import torch
    import torch.distributed as dist

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(value: float, world: int, local: int) -> float:
    import torch
    import torch.distributed as dist

    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(values, world: int, local: int):
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def fixed_job(config: str, total_spp: int, small: bool, world: int, rank: int, local: int, pool: int = 0, seed: int = 4242, check_single: bool = False):
    """A FIXED job (strong scaling): `total_spp` sample passes of `config` split over the ranks, timed from the project IR in
    host memory to the developed image in host memory on rank 0, with the phases the job consists of.  Every phase is the
    max over ranks (barrier + wall clock around blocking library calls; `render` is also reported in device time)."""
    import torch

    from pyrite_b200 import api
    from pyrite_b200.distributed import init_film_comm, shard_for_rank

    ir = build_project(config, small)   # building the stand-in mesh is test-data generation, not part of the job
    phases = {}
    verbose = os.environ.get("PYR_BENCH_VERBOSE") and rank == 0

    def note(what):
        if verbose:
            print(f"[fixed_job {config} {total_spp} spp] {what} at {time.time() - t_job:.2f} s", file=sys.stderr, flush=True)

    barrier(world)
    t_job = time.time()
    t0 = time.time()
    r = api.Renderer(local)
    r.load(ir)                          # IR decode + scene build (host BVH build, replicated on every rank) + upload
    torch.cuda.synchronize()
    phases["load_build_upload"] = max_over_ranks(time.time() - t0, world, local)
    note("loaded")
    t0 = time.time()
    if world > 1:
        init_film_comm(r, wait=False)   # pyr_comm_init_async: NCCL's set-up runs under the render, pyr_film_reduce waits for it
    phases["comm_init"] = max_over_ranks(time.time() - t0, world, local)
    offset, stride = shard_for_rank(rank, world)
    r.counters(reset=True)
    t0 = time.time()
    device_s = r.render(seed=seed, spp=total_spp, sample_offset=offset, sample_stride=stride, pool_paths=pool)   # the production path: two wavefront lanes
    phases["render"] = max_over_ranks(time.time() - t0, world, local)
    render_device = max_over_ranks(device_s, world, local)
    note("rendered")
    c = r.counters()
    t0 = time.time()
    if world > 1:
        r.film_reduce(0)
    phases["film_reduce"] = max_over_ranks(time.time() - t0, world, local)
    t0 = time.time()
    xyz = None
    if rank == 0:
        xyz, srgb = r.develop()
    phases["develop_download"] = max_over_ranks(time.time() - t0, world, local)
    note("developed")
    total_s = max_over_ranks(time.time() - t_job, world, local)
    rays, samples = sum_over_ranks([c["rays"], c["path_samples"]], world, local)
    info = r.info
    out = {
        "config": workload_name(config, small), "job_spp": total_spp, "n_gpus": world, "seconds_total": total_s, "seconds": phases,
        "render_device_seconds": render_device, "rays": rays, "path_samples": samples, "kernel_launches": int(c["kernel_launches"]),
        "mrays_per_s_render": rays / render_device / 1e6, "mrays_per_s_job": rays / total_s / 1e6,
        "msamples_per_s_render": samples / render_device / 1e6, "msamples_per_s_job": samples / total_s / 1e6,
        "film_bytes": int(info.width) * int(info.height) * int(info.bins) * 8, "ir_bytes": len(ir), "image_bytes": int(info.width) * int(info.height) * 15,
        "limiter": max(phases, key=phases.get),
    }
    if config == "C4" and rank == 0:
        out["sphere_tracing_roofline"] = sphere_tracing_roofline(r, total_spp, pool)
        note("sphere-tracing roofline passes done")
    if check_single and world > 1:
        # the sharded job must be THE single-GPU job: rank 0 renders a small job alone and the ranks render it together
        r.render(seed=seed + 1, spp=2 * world, sample_offset=offset, sample_stride=stride, pool_paths=pool)
        r.film_reduce(0)
        if rank == 0:
            together = r.film()
            r.render(seed=seed + 1, spp=2 * world, pool_paths=pool)
            alone = r.film()
            out["sharding_check"] = {"weights_equal": bool(np.array_equal(together[..., 1], alone[..., 1])) if config != "C5" else
                                     bool(np.allclose(together[..., 1], alone[..., 1], rtol=2e-4, atol=1e-5)),
                                     "accumulators_close": bool(np.allclose(together[..., 0], alone[..., 0], rtol=2e-4, atol=1e-5))}
        barrier(world)
    if rank == 0 and xyz is not None:
        # a sample whose lamp-vertex BRDF is 0 exposes 0/0 (the reference's `brdf_in = x / x`, bidirectional.rs:365-372): such bins stay NaN there too
        finite = np.isfinite(xyz[..., 1])
        out["mean_luminance"] = float(xyz[..., 1][finite].mean()) if finite.any() else None
        out["non_finite_pixels"] = int((~finite).sum())
    r.close()
    return out


def sphere_tracing_roofline(r, job_spp: int, pool: int) -> dict:
    """FP32 roofline of the sphere-tracing stage: distance-estimator iterations counted by the kernel's own
    statistics variant on one sample pass, times the operation counts of one iteration of `Mandelbulb::get` (30 FP32 + 12
    SFU-class, distance_estimators.rs:20-36) resp. cubic `QuaternionJulia::get` (128 FP32 + 1 sqrt, :59-65,82,90), over the
    device time of the traversal + sphere-tracing launches of a one-wavefront timing pass, against the measured FFMA peak."""
    hbm, hbm_src, ffma, ffma_src = measured_peaks()
    r.counters(reset=True)
    r.render(seed=9, spp=1, pool_paths=pool, stats=True)
    cs = r.counters()
    r.counters(reset=True)
    r.render(seed=9, spp=1, pool_paths=pool, timing=True)
    ct = r.counters()
    julia, bulb = cs["julia_iterations"], cs["march_iterations"] - cs["julia_iterations"]
    flops = bulb * (30 + 12) + julia * (128 + 1)
    secs = ct["trace_seconds"]
    achieved = flops / max(secs, 1e-12) / 1e12
    return {"bound": "fp32", "kernel": "k_trace + k_march (+ k_march_apply)", "achieved": achieved, "peak": ffma, "unit": "TFLOP/s", "frac": achieved / ffma,
            "peak_source": ffma_src, "de_evaluations_per_ray": cs["de_evals"] / max(cs["rays"], 1), "iterations_per_de_evaluation": cs["de_iterations"] / max(cs["de_evals"], 1), "rays_of_the_pass": cs["rays"],
            "mandelbulb_iterations": bulb, "julia_iterations": julia, "ops_per_iteration": {"mandelbulb": 42, "julia_cubic": 129},
            "seconds_of_one_pass": secs, "share_of_render": secs / max(ct["trace_seconds"] + ct["shade_seconds"], 1e-12),
            "note": "the time is that of the whole traversal stage of a one-wavefront pass (BVH walk + sphere tracing); profiles/ has the sphere-tracing kernel's share and its pipe utilisation"}


def run_fixed(args):
    """--config CX: one line for a fixed job (strong scaling)."""
    world, rank, local = dist_setup()
    config = args.config
    spp = args.spp or CONFIGS[config]["spp"]
    sampler = ClockSampler(local)
    for _ in range(max(args.warmup, 1) if not args.small else 1):   # warm-up: a 1/16 job (library load, allocator, NCCL)
        fixed_job(config, max(spp // 16, world), args.small, world, rank, local, args.pool)
    if rank == 0:
        sampler.start()
    t0 = time.time()
    job = fixed_job(config, spp, args.small, world, rank, local, args.pool, check_single=(config != "C5" or args.small))
    t1 = time.time()
    if rank == 0:
        clocks = sampler.stop(t0, t1)
        hbm, hbm_src, ffma, ffma_src = measured_peaks()
        line = {
            "metric": "Mrays/s", "value": job["mrays_per_s_job"], "unit": "Mrays/s", "path_samples_per_s": job["msamples_per_s_job"] * 1e6,
            "n_gpus": world, "steps": 1, "warmup": args.warmup, "ms_per_step": 1e3 * job["seconds_total"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, world), "clocks": clocks,
            "e2e": {"value": job["mrays_per_s_job"], "unit": "Mrays/s", "h2d_bytes_per_step": job["ir_bytes"], "d2h_bytes_per_step": job["image_bytes"],
                    "what": "the whole job through the public API: project IR in host memory -> pyr_project_load -> pyr_render -> pyr_film_reduce -> pyr_film_develop -> images in host memory, wall clock"},
            "gpu_launches": job.get("kernel_launches"), "roofline": job.get("sphere_tracing_roofline"), "fixed_job": job,
        }
        print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    import torch.distributed as dist

    if world > 1 and dist.is_initialized():
        dist.destroy_process_group()


def run_product(args):
    import torch

    from pyrite_b200 import api
    from pyrite_b200.distributed import init_film_comm, shard_for_rank

    world, rank, local = dist_setup()
    ir = build_project("C2", args.small)
    r = api.Renderer(local)
    r.load(ir)
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    if world > 1:
        init_film_comm(r)
    info = r.info
    spp_step = args.spp_per_step
    offset, stride = shard_for_rank(rank, world)

    def step(k, **kw):
        # every rank renders its own sample passes: indices offset, offset+stride, ... of a (world * spp_step)-spp job
        return r.render(seed=1000 + k, spp=spp_step * world, sample_offset=offset, sample_stride=stride, reset_film=(k == 0), pool_paths=args.pool, **kw)

    for w in range(args.warmup):
        step(w)
    if world > 1:
        r.film_reduce(0)   # warm-up of the collective too (NCCL sizes its buffers at the first large reduction)
    r.counters(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier(world)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record(stream)
    for k in range(args.steps):
        step(k)           # the production path: two wavefront lanes, no per-kernel events
    if world > 1:
        r.film_reduce(0)  # the one NCCL film reduction of the job (pyr_film_reduce, on the same stream)
    ev1.record(stream)
    barrier(world)
    t1 = time.time()
    device_s = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3, world, local)
    c = r.counters()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    total_rays, total_samples, launches = sum_over_ranks([c["rays"], c["path_samples"], c["kernel_launches"]], world, local)
    # the same steps once more in the one-lane timing mode: CUDA events around every trace / shade launch, for the rooflines
    if rank == 0:
        step(0, timing=True)   # untimed: the one-lane pool is allocated here
        r.counters(reset=True)
        for k in range(args.steps):
            step(k, timing=True)
        c = r.counters()

    # ---- end-to-end through the public API with host buffers: render step + develop + image download
    e2e_steps = max(1, min(args.steps, 4))
    if rank == 0:  # the images land in page-locked host memory
        host_xyz = torch.empty((info.height, info.width, 3), dtype=torch.float32, pin_memory=True).numpy()
        host_srgb = torch.empty((info.height, info.width, 3), dtype=torch.uint8, pin_memory=True).numpy()
        r.develop(out_xyz=host_xyz, out_srgb=host_srgb)  # untimed: one-off white-balance scan and first use of the develop kernels
    barrier(world)
    r.counters(reset=True)
    te0 = time.time()
    for k in range(e2e_steps):
        step(k)
        if world > 1:
            r.film_reduce(0)
        if rank == 0:
            r.develop(out_xyz=host_xyz, out_srgb=host_srgb)
    barrier(world)
    te1 = time.time()
    e2e_mrays = sum_over_ranks([r.counters()["rays"]], world, local)[0] / (te1 - te0) / 1e6
    pixels = info.width * info.height

    roofline = cpu = None
    if rank == 0:
        hbm, hbm_src, ffma, ffma_src = measured_peaks()
        # ---- the traversal kernel: its own fetch counts (STATS variant) on one pass of the same workload
        r.counters(reset=True)
        r.render(seed=1000, spp=1, sample_offset=0, sample_stride=1, reset_film=False, pool_paths=args.pool, stats=True)
        cs = r.counters()
        nodes_per_ray = cs["node_fetches"] / max(cs["rays"], 1)
        boxes_per_ray = cs["nodes_visited"] / max(cs["rays"], 1)
        leaves_per_ray = cs["leaves_tested"] / max(cs["rays"], 1)
        own_bytes_per_ray = RAY_BYTES + HIT_BYTES + nodes_per_ray * NODE_BYTES + leaves_per_ray * PRIM_BYTES
        # ---- CPU baseline: the oracle port on a bounded sample of the same workload; it also counts the reference-order walk
        ref_nodes = ref_leaves = None
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            variant, how = oracle_variant()
            secs, oc = oracle_leg(ir, threads, args.cpu_spp, args.cpu_fraction, variant=variant)
            ref_nodes, ref_leaves = oc["nodes"] / max(oc["rays"], 1), oc["leaves"] / max(oc["rays"], 1)
            cpu = {"value": oc["rays"] / secs / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port", "build": how,
                   "sample": f"{args.cpu_spp}/{args.cpu_fraction} sample passes over the whole C2 image ({oc['path_samples']} path samples, {oc['rays']} rays, {secs:.1f} s)",
                   "path_samples_per_s": oc["path_samples"] / secs}
            if variant == "native":  # the portable build the parity tests use, for comparison
                secs_p, oc_p = oracle_leg(ir, threads, 1, args.cpu_fraction, variant="glibc")
                cpu["value_portable_O2_build"] = oc_p["rays"] / secs_p / 1e6
        traffic = {}
        tfile = ROOT / "profiles" / "kernel_traffic.json"
        if tfile.exists():
            try:
                traffic = json.loads(tfile.read_text())
            except Exception:
                traffic = {}
        trace_s, trace_n = c["trace_seconds"], max(c["trace_launches"], 1)
        shade_s, shade_n = c["shade_seconds"], max(c["shade_launches"], 1)
        rays_per_launch = c["rays"] / trace_n
        launch_s = trace_s / trace_n
        # the ALGORITHMIC figure is the reference-order walk (binary tree, pre-order, 1 triangle per leaf) counted by the
        # oracle on the same sample mix; the kernel's own (smaller) fetch volume and its real DRAM / L2 traffic are given next to it
        ref_bytes_per_ray = (REF_RAY_IO + ref_nodes * REF_NODE_BYTES + ref_leaves * REF_TRI_BYTES) if ref_nodes is not None else None
        bytes_per_ray = ref_bytes_per_ray if ref_bytes_per_ray is not None else own_bytes_per_ray
        achieved = rays_per_launch * bytes_per_ray / max(launch_s, 1e-12) / 1e9
        # measured DRAM / L2 bytes per ray of a whole profiled render (tools/summarize_ncu.py --traffic), scaled to this run's launches
        t_dram = traffic["k_trace_dram_bytes_per_ray"] * rays_per_launch if "k_trace_dram_bytes_per_ray" in traffic else traffic.get("k_trace")
        t_l2 = traffic["k_trace_l2_bytes_per_ray"] * rays_per_launch if "k_trace_l2_bytes_per_ray" in traffic else traffic.get("k_trace_l2_bytes")
        trace = {
            "bound": "hbm", "kernel": "k_trace", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": t_dram,
            "peak_source": hbm_src, "bytes_per_ray": bytes_per_ray,
            "bytes_per_ray_is": ("SURVEY.md §8(d): 52 + Vn*32 + Vt*48 with the reference-order counts of the oracle on the same workload" if ref_bytes_per_ray is not None
                                 else "the kernel's own fetches (no CPU leg ran to count the reference-order walk)"),
            "reference_order_nodes_per_ray": ref_nodes, "reference_order_leaves_per_ray": ref_leaves,
            "own_bytes_per_ray": own_bytes_per_ray, "own_nodes_fetched_per_ray": nodes_per_ray, "own_boxes_tested_per_ray": boxes_per_ray, "own_leaves_per_ray": leaves_per_ray,
            "achieved_own": rays_per_launch * own_bytes_per_ray / max(launch_s, 1e-12) / 1e9,
            "frac_own": rays_per_launch * own_bytes_per_ray / max(launch_s, 1e-12) / 1e9 / hbm,
            "dram_frac": (t_dram / max(launch_s, 1e-12) / 1e9 / hbm) if t_dram else None,
            "l2_gbs": (t_l2 / max(launch_s, 1e-12) / 1e9) if t_l2 else None,
            "limited_by": traffic.get("k_trace_limited_by", "instruction issue (see profiles/): the walk is served by L1/L2, DRAM is nearly idle"),
            "rays_per_launch": rays_per_launch, "avg_launch_ms": 1e3 * launch_s, "share_of_step": trace_s / max(trace_s + shade_s, 1e-12),
            "mrays_per_s": c["rays"] / max(trace_s, 1e-12) / 1e6,
            "note": "frac > 1 is possible: the kernel visits fewer boxes than the reference-order walk and caches serve most fetches; dram_frac is the real DRAM utilisation (traffic from the ncu capture)",
        }
        # the shade stage (k_bin_* + k_wave_simple): per path iteration the path record (64 B header + ceil(3 S / 8) 32-byte chunks) is read and
        # written, the path ray and its hit are read, the next ray is written; per visibility ray 4 B result in, 32 B ray + 32 B pending light
        # out and the light back in; per sample S film atomics of 8 B
        S = info.spectrum_samples
        core_bytes = 64 + 32 * ((3 * S + 7) // 8)
        path_iterations = c["path_rays"] + c["path_samples"]
        shadow = c["rays"] - c["path_rays"]
        shade_bytes = path_iterations * 2 * core_bytes + c["path_rays"] * (32 + 32 + 32) + shadow * (4 + 32 + 64) + c["path_samples"] * S * 8
        shade_achieved = shade_bytes / max(shade_s, 1e-12) / 1e9
        t_shade = (traffic["shade_dram_bytes_per_path_iteration"] * path_iterations / shade_n if "shade_dram_bytes_per_path_iteration" in traffic
                   else traffic.get("k_wave_simple"))
        shade = {"bound": "hbm", "kernel": "k_bin_* + k_wave_simple", "achieved": shade_achieved, "peak": hbm, "unit": "GB/s", "frac": shade_achieved / hbm,
                 "traffic": t_shade, "peak_source": hbm_src, "bytes_per_path_iteration": shade_bytes / max(path_iterations, 1),
                 "dram_frac": (t_shade / max(shade_s / shade_n, 1e-12) / 1e9 / hbm) if t_shade else None,
                 "limited_by": traffic.get("k_wave_simple_limited_by", "memory latency at low occupancy (see profiles/)"),
                 "path_iterations_per_launch": path_iterations / shade_n, "avg_launch_ms": 1e3 * shade_s / shade_n,
                 "share_of_step": shade_s / max(trace_s + shade_s, 1e-12)}
        roofline = dict(shade if shade_s > trace_s else trace)   # the dominant kernel by measured device time
        roofline["other_kernel"] = trace if shade_s > trace_s else shade
        roofline["fp32_peak_tflops"] = ffma
        roofline["fp32_peak_source"] = ffma_src
    r.close()

    # ---- a small FIXED job of the north_star target's kind (C5: 4K, bidirectional), split over the ranks: strong scaling
    strong = None
    if args.strong_spp > 0:
        fixed_job("C5", max(args.strong_spp // 8, world), args.small, world, rank, local)   # warm-up
        strong = fixed_job("C5", max(args.strong_spp, world), args.small, world, rank, local, check_single=False)
        check = fixed_job("C2", 4 * world, True, world, rank, local, check_single=True)        # sharded == single-GPU, on a small scene
        if rank == 0:
            strong["sharding_check_on_small_C2"] = check.get("sharding_check")
            strong["note"] = ("a FIXED job split over the ranks, timed from the IR in host memory to the developed image on rank 0; the per-rank fixed costs "
                              "(scene build and upload in load_build_upload, develop) do not shrink with N, so a job a quarter of the target's length scales a "
                              "little worse than the 1024-spp target itself (profiles/ holds that run); the communicator is set up under the render "
                              "(pyr_comm_init_async), film_reduce includes whatever wait for it is left")

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": total_rays / device_s / 1e6, "unit": "Mrays/s",
            "path_samples_per_s": total_samples / device_s,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * device_s / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, world),
            "pool_paths": args.pool or "library default: 2^24 paths in flight (capped by a sixth of device memory and by the step's path samples)",
            "clocks": clocks,
            "e2e": {"value": e2e_mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 32, "d2h_bytes_per_step": pixels * 15,
                    "what": "pyr_render step + pyr_film_reduce + pyr_film_develop + XYZ f32 and sRGB u8 image download to pinned host memory, wall clock"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "strong_scaling": strong,
            "wall_s_timed_region": t1 - t0,
        }
        print(json.dumps(line), flush=True)
    finish(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="run ONE fixed job of this BASELINE config (strong scaling) instead of the C2 step benchmark")
    ap.add_argument("--spp", type=int, default=0, help="--config: total sample passes of the fixed job (0 = the config's own: C3 200, C4 512, C5 1024)")
    ap.add_argument("--spp-per-step", type=int, default=32, help="sample passes per GPU per step (8 steps of 32 = the 256-spp C2 job)")
    ap.add_argument("--strong-spp", type=int, default=256, help="sample passes of the fixed C5 job reported under strong_scaling: a quarter of the 1024-spp target (0 = skip)")
    ap.add_argument("--pool", type=int, default=0, help="paths in flight (0 = library default)")
    ap.add_argument("--cpu-fraction", type=int, default=1, help="the CPU legs render every N-th path sample of a pass")
    ap.add_argument("--cpu-spp", type=int, default=4, help="sample passes of the cpu_baseline leg (4 passes = about 15 s on 16 cores)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--small", action="store_true", help="tiny stand-in workload for plumbing tests (not a bench result)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "product" and not args.small and not args.config:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.config:
        run_fixed(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
