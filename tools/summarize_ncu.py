#!/usr/bin/env python3
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_ncu.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof.ncu-rep]

writes profiles/<tag>_launches.md (per-kernel totals and shares from the gpu__time_duration launch list),
profiles/<tag>_kernels.md (key metrics of every kernel captured with --set full) and updates
profiles/kernel_traffic.json (DRAM bytes per launch of k_trace and k_wave_simple, read by bench.py as roofline.traffic).

    python tools/summarize_ncu.py <tag> --traffic gpurun_out/traffic.csv --counts gpurun_out/traffic_counts.json

is the better source of that file: `traffic.csv` lists dram__bytes_read/write.sum, lts__t_bytes.sum and gpu__time_duration.sum of EVERY
launch of one whole render (tools/profile_step.py with PYR_NO_WARMUP=1 PYR_COUNTS_JSON=...), `traffic_counts.json` holds that render's ray
and path-iteration counts, and the result is DRAM / L2 bytes PER RAY (k_trace) and PER PATH ITERATION (k_bin_* + k_wave_simple), which bench.py
scales to the rays of its own launches - a single captured launch may be a small one from the drain tail of a render."""
import argparse
import csv
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PROF = ROOT / "profiles"

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
]


def short(name):
    name = name.split("(")[0]
    return name.split("::")[-1].strip()


def launches(path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        tot[short(r[ki])] += float(r[vi].replace(",", "")) / 1e6
        cnt[short(r[ki])] += 1
    total = sum(tot.values())
    out = [f"# {tag}: launch list (ncu --metrics gpu__time_duration.sum --clock-control none)", "",
           "Per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes.", "",
           "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append(f"| {k} | {cnt[k]} | {v:.3f} | {100 * v / total:.1f}% |")
    out.append(f"| **all** | {sum(cnt.values())} | {total:.3f} | 100% |")
    (PROF / f"{tag}_launches.md").write_text("\n".join(out) + "\n")
    print("\n".join(out))


def kernels(rep, tag, keep_traffic=False):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(list)
    for r in rows[2:]:
        per[short(r[col["Kernel Name"]])].append(r)
    out = [f"# {tag}: per-kernel metrics (ncu --set full --clock-control none --import-source on)", ""]
    traffic = {}
    for k, rs in per.items():
        out += [f"## {k} ({len(rs)} launch(es) captured; mean over them)", "", "| metric | value | unit |", "|---|---:|---|"]
        for m in METRICS:
            if m not in col:
                continue
            vals = []
            for r in rs:
                try:
                    vals.append(float(r[col[m]].replace(",", "")))
                except ValueError:
                    pass
            if vals:
                out.append(f"| {m} | {sum(vals) / len(vals):,.3f} | {units[col[m]]} |")
        def mean(m):
            v = [float(r[col[m]].replace(",", "")) for r in rs]
            return sum(v) / len(v)
        unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = mean("dram__bytes_read.sum") * unit_scale.get(units[col["dram__bytes_read.sum"]], 1)
        wr = mean("dram__bytes_write.sum") * unit_scale.get(units[col["dram__bytes_write.sum"]], 1)
        traffic[k] = rd + wr
        out += [f"| dram bytes per launch (read + write) | {rd + wr:,.0f} | byte |", ""]
    (PROF / f"{tag}_kernels.md").write_text("\n".join(out) + "\n")
    print("\n".join(out))
    keep = {}
    for k, v in traffic.items():
        if k.startswith("k_trace"):
            keep["k_trace"] = v
        elif k.startswith("k_wave_simple"):
            keep["k_wave_simple"] = v
    if keep and not keep_traffic:
        keep["source"] = f"profiles/{tag}_kernels.md (dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full)"
        (PROF / "kernel_traffic.json").write_text(json.dumps(keep) + "\n")


def traffic(path, counts_path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    tot = defaultdict(lambda: defaultdict(float))
    n = defaultdict(int)
    for r in rows[1:]:
        k = short(r[ki])
        tot[k][r[mi]] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
        if r[mi] == "gpu__time_duration.sum":
            n[k] += 1
    counts = json.loads(Path(counts_path).read_text())
    rays, path_iterations = counts["rays"], counts["path_rays"] + counts["path_samples"]
    def group(pred):
        g = defaultdict(float)
        for k, m in tot.items():
            if pred(k):
                for name, v in m.items():
                    g[name] += v
        return g
    tr, sh = group(lambda k: k.startswith("k_trace")), group(lambda k: k.startswith("k_wave_simple") or k.startswith("k_bin"))
    out = {
        "k_trace_dram_bytes_per_ray": (tr["dram__bytes_read.sum"] + tr["dram__bytes_write.sum"]) / rays,
        "k_trace_l2_bytes_per_ray": tr["lts__t_bytes.sum"] / rays,
        "k_trace_seconds_under_ncu": tr["gpu__time_duration.sum"],
        "shade_dram_bytes_per_path_iteration": (sh["dram__bytes_read.sum"] + sh["dram__bytes_write.sum"]) / path_iterations,
        "shade_l2_bytes_per_path_iteration": sh["lts__t_bytes.sum"] / path_iterations,
        "shade_seconds_under_ncu": sh["gpu__time_duration.sum"],
        "run": counts,
        "launches": {k: n[k] for k in sorted(n)},
        "source": f"whole-run ncu metric pass ({Path(path).name}: dram__bytes_read.sum + dram__bytes_write.sum, lts__t_bytes.sum of every launch of one render) "
                  f"divided by the render's own counts; {tag}",
    }
    (PROF / "kernel_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--traffic")
    ap.add_argument("--counts")
    ap.add_argument("--keep-traffic", action="store_true", help="--rep: do not overwrite profiles/kernel_traffic.json")
    a = ap.parse_args()
    PROF.mkdir(exist_ok=True)
    if a.launches:
        launches(a.launches, a.tag)
    if a.rep:
        kernels(a.rep, a.tag, a.keep_traffic)
    if a.traffic:
        traffic(a.traffic, a.counts, a.tag)


if __name__ == "__main__":
    main()
