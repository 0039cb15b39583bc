#!/usr/bin/env python3
"""Benchmark of the render hot path (BASELINE.json metric: Mrays/s and spectral path samples/s).

Workload: BASELINE config C2 - the ~870k-triangle dragon stand-in, 1920x1080, camera-to-light
integrator with S=10 wavelengths, B=8 bounces, L=4 light samples.  One *step* = `--spp-per-step`
sample passes over the whole image (256 / spp-per-step steps make the full 256-spp config), so
throughput does not depend on how many steps are timed.  A ray = one World::intersect call, a path
sample = one render_tile iteration.

    python bench.py --gpus 1 --steps 8 --warmup 3            # product arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps 3 --warmup 1    # CPU arm: the oracle port on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md §8 for the definitions of every field.
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

SCENE = dict(width=1920, height=1080, spp=256, bounces=8, light_samples=4, spectrum_samples=10)
WORKLOAD = "C2 dragon stand-in 871,200 tris + 3 planes + point light, 1920x1080, simple integrator S=10 B=8 L=4"
NODE_BYTES, PRIM_BYTES, RAY_BYTES, HIT_BYTES = 128, 48, 32, 32  # Node4, Prim, Ray, Hit records (device_types.h)


def build_project(args):
    from pyrite_b200 import project, scenes

    kw = dict(SCENE)
    if args.small:
        kw.update(width=480, height=270)
        kw["mesh"] = scenes.dragon_mesh(400, 50)
    return project.serialize_project(scenes.dragon(**kw))


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


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def oracle_leg(ir: bytes, fraction_den: int, threads: int, seed: int = 1, spp: int = 1):
    """The CPU restatement of pyrite's renderer (oracle/) on `spp`/fraction_den sample passes."""
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_lib import Oracle

    o = Oracle(ir)
    o.counters(reset=True)
    secs = o.render(seed=seed, spp=spp, sample_offset=0, sample_stride=fraction_den, threads=threads, cas_attempts=5)
    c = o.counters()
    return o, secs, c


def run_reference(args):
    """--impl reference: pyrite's own CPU algorithm (the oracle port; the Rust crate cannot be built
    in this image) with all host threads, each step a bounded sample of the C2 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_lib import Oracle

    threads = os.cpu_count() or 1
    ir = build_project(args)
    o = Oracle(ir)
    den = args.cpu_fraction
    for w in range(args.warmup):
        o.render(seed=100 + w, spp=1, sample_offset=w % den, sample_stride=den * 4, threads=threads, cas_attempts=5)
    o.counters(reset=True)
    total = 0.0
    for k in range(args.steps):
        total += o.render(seed=k, spp=1, sample_offset=k % den, sample_stride=den, threads=threads, cas_attempts=5, reset_film=(k == 0))
    c = o.counters()
    mrays = c["rays"] / total / 1e6
    sample = f"{args.steps} steps x 1/{den} of one sample pass over the whole C2 image ({c['path_samples']} path samples, {c['rays']} rays, {total:.1f} s)"
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "path_samples_per_s": c["path_samples"] / total,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD if not args.small else WORKLOAD + " [--small: 480x270, 40k tris]", "host": "CPU only"},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_product(args):
    import torch
    import torch.distributed as dist

    from pyrite_b200 import api
    from pyrite_b200.distributed import reduce_film, shard_for_rank

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    ir = build_project(args)
    r = api.Renderer(local)
    r.load(ir)
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    info = r.info
    spp_step = args.spp_per_step
    offset, stride = shard_for_rank(rank, world)

    def step(k, **kw):
        # every rank renders its own sample passes: indices offset, offset+stride, ... of a (world * spp_step)-spp job
        return r.render(seed=1000 + k, spp=spp_step * world, sample_offset=offset, sample_stride=stride, reset_film=(k == 0), pool_paths=args.pool, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(w)
    r.counters(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record(stream)
    for k in range(args.steps):
        step(k, timing=True)
    film = torch.as_tensor(r.film_device(), device=f"cuda:{local}")
    if world > 1:
        reduce_film(film, dst=0)  # the one NCCL film reduction of the job
    ev1.record(stream)
    barrier()
    t1 = time.time()
    device_s = ev0.elapsed_time(ev1) * 1e-3
    c = r.counters()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    stats = torch.tensor([device_s, float(c["rays"]), float(c["path_samples"]), float(c["kernel_launches"])], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        tmax = stats.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        device_s = float(tmax[0])
    total_rays, total_samples, launches = float(stats[1]), float(stats[2]), int(stats[3])

    # ---- end-to-end through the public API with host buffers: render step + develop + image download
    e2e_steps = max(1, min(args.steps, 4))
    if rank == 0:  # the images land in page-locked host memory
        host_xyz = torch.empty((info.height, info.width, 3), dtype=torch.float32, pin_memory=True).numpy()
        host_srgb = torch.empty((info.height, info.width, 3), dtype=torch.uint8, pin_memory=True).numpy()
        r.develop(out_xyz=host_xyz, out_srgb=host_srgb)  # untimed: one-off white-balance scan and first use of the develop kernels
    barrier()
    r.counters(reset=True)
    te0 = time.time()
    for k in range(e2e_steps):
        step(k)
        if world > 1:
            reduce_film(film, dst=0)
        if rank == 0:
            xyz, srgb = r.develop(out_xyz=host_xyz, out_srgb=host_srgb)
    barrier()
    te1 = time.time()
    ce = r.counters()
    e2e = torch.tensor([float(ce["rays"])], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e, op=dist.ReduceOp.SUM)
    e2e_mrays = float(e2e[0]) / (te1 - te0) / 1e6
    pixels = info.width * info.height

    if rank == 0:
        # ---- rooflines of the two kernels of the step: traversal (algorithmic bytes from its own STATS counters) and shade
        r.counters(reset=True)
        r.render(seed=1000, spp=1, sample_offset=0, sample_stride=1, reset_film=False, pool_paths=args.pool, stats=True)
        cs = r.counters()
        nodes_per_ray = cs["node_fetches"] / max(cs["rays"], 1)
        boxes_per_ray = cs["nodes_visited"] / max(cs["rays"], 1)
        leaves_per_ray = cs["leaves_tested"] / max(cs["rays"], 1)
        bytes_per_ray = RAY_BYTES + HIT_BYTES + nodes_per_ray * NODE_BYTES + leaves_per_ray * PRIM_BYTES
        trace_s, trace_n = c["trace_seconds"], max(c["trace_launches"], 1)
        rays_per_launch = c["rays"] / trace_n
        achieved = (c["rays"] * bytes_per_ray) / max(trace_s, 1e-12) / 1e9
        peak, peak_src = measured_peak_hbm()
        traffic = {}
        tfile = ROOT / "profiles" / "kernel_traffic.json"
        if tfile.exists():
            try:
                traffic = json.loads(tfile.read_text())
            except Exception:
                traffic = {}
        trace = {"bound": "hbm", "kernel": "k_trace", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                 "traffic": traffic.get("k_trace"), "peak_source": peak_src, "bytes_per_ray": bytes_per_ray, "nodes_fetched_per_ray": nodes_per_ray,
                 "boxes_tested_per_ray": boxes_per_ray, "leaves_per_ray": leaves_per_ray, "rays_per_launch": rays_per_launch,
                 "avg_launch_ms": 1e3 * trace_s / trace_n, "share_of_step": trace_s / max(c["trace_seconds"] + c["shade_seconds"], 1e-12),
                 "mrays_per_s": c["rays"] / max(trace_s, 1e-12) / 1e6,
                 "note": "algorithmic node/primitive bytes; the ncu capture shows that L2/L1 serve most of them (traffic = DRAM bytes per launch)"}
        # the shade stage (k_bin_keys / _scan / _scatter + k_wave_simple): per path iteration the path record (64 B header + ceil(3 S / 8) 32-byte chunks of
        # wavelengths / brightness / reflectance) is read and written, the path ray and its hit are read, the next ray is written; per
        # visibility ray 4 B result in, 32 B ray + 32 B pending light out and the light back in; per sample S film atomics of 8 B
        S = info.spectrum_samples
        core_bytes = 64 + 32 * ((3 * S + 7) // 8)
        path_iterations = c["path_rays"] + c["path_samples"]
        shadow = c["rays"] - c["path_rays"]
        shade_bytes = path_iterations * 2 * core_bytes + c["path_rays"] * (32 + 32 + 32) + shadow * (4 + 32 + 64) + c["path_samples"] * S * 8
        shade_s, shade_n = c["shade_seconds"], max(c["shade_launches"], 1)
        shade_achieved = shade_bytes / max(shade_s, 1e-12) / 1e9
        shade = {"bound": "hbm", "kernel": "k_bin_* + k_wave_simple", "achieved": shade_achieved, "peak": peak, "unit": "GB/s", "frac": shade_achieved / peak,
                 "traffic": traffic.get("k_wave_simple"), "peak_source": peak_src, "bytes_per_path_iteration": shade_bytes / max(path_iterations, 1),
                 "path_iterations_per_launch": path_iterations / shade_n, "avg_launch_ms": 1e3 * shade_s / shade_n,
                 "share_of_step": shade_s / max(c["trace_seconds"] + c["shade_seconds"], 1e-12)}
        roofline = dict(shade if shade_s > trace_s else trace)   # the dominant kernel by measured device time
        roofline["other_kernel"] = trace if shade_s > trace_s else shade
        # ---- CPU baseline: the oracle port on a bounded sample of the same workload
        cpu = None
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            o, secs, oc = oracle_leg(ir, args.cpu_fraction, threads, spp=args.cpu_spp)
            cpu = {"value": oc["rays"] / secs / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_spp}/{args.cpu_fraction} sample passes over the whole C2 image ({oc['path_samples']} path samples, {oc['rays']} rays, {secs:.1f} s)",
                   "path_samples_per_s": oc["path_samples"] / secs}
        line = {
            "metric": "Mrays/s", "value": total_rays / device_s / 1e6, "unit": "Mrays/s",
            "path_samples_per_s": total_samples / device_s,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * device_s / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if not args.small else WORKLOAD + " [--small: 480x270, 40k tris]",
                       "step": f"{spp_step} spp per GPU per step ({256 // max(spp_step, 1)} steps = the 256-spp config on 1 GPU)",
                       "parallelism": f"sample-pass sharding over {world} GPU(s), one NCCL film reduce at the end",
                       "l2": "working set (film 1.06 GB + path pool + 150 MB BVH) exceeds the 126 MB L2; no flush needed",
                       "pool_paths": args.pool or "library default: 2^24 paths in flight (capped by a sixth of device memory and by the step's path samples)"},
            "clocks": clocks,
            "e2e": {"value": e2e_mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 32, "d2h_bytes_per_step": pixels * 15,
                    "what": "pyr_render step + pyr_film_develop + XYZ f32 and sRGB u8 image download to pinned host memory, wall clock"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "wall_s_timed_region": t1 - t0,
        }
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=32, help="sample passes per GPU per step (8 steps of 32 = the 256-spp C2 job)")
    ap.add_argument("--pool", type=int, default=0, help="paths in flight (0 = library default)")
    ap.add_argument("--cpu-fraction", type=int, default=1, help="the CPU legs render every N-th path sample of a pass")
    ap.add_argument("--cpu-spp", type=int, default=4, help="sample passes of the cpu_baseline leg (4 passes = about 15 s on 16 cores)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--small", action="store_true", help="tiny stand-in workload for plumbing tests (not a bench result)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "product":
        args.warmup = max(args.warmup, 3) if not args.small else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
