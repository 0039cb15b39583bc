#!/usr/bin/env python3
"""Hot source lines of a kernel from an ncu report captured with --import-source on (needs -lineinfo):
    python tools/ncu_hot_lines.py report.ncu-rep <kernel regex> [top N] [launch id]
Aggregates the warp-stall samples of `ncu --page source --print-source sass,cuda` per CUDA source line and prints the top N
with their dominant stall reasons."""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv", "--kernel-name", f"regex:{kernel}"]
if len(sys.argv) > 4:
    cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
text = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
file = None
hdr = None
lines = {}
first_kernel = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if first_kernel is None:
            first_kernel = r[1]
        cur_kernel = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or cur_kernel != first_kernel:
        continue
    if r[2] != "-":   # a SASS row
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        samples = int(d["# Samples"])
    except Exception:
        continue
    if samples == 0:
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "(Not Issued)" not in k and v.isdigit() and int(v) > 0}
    key = (file, int(r[0]))
    if key in lines:
        lines[key][1] += samples
        for k, v in stalls.items():
            lines[key][2][k] = lines[key][2].get(k, 0) + v
        lines[key][3] += int(d.get("Instructions Executed", "0") or 0)
    else:
        lines[key] = [r[1].strip()[:110], samples, stalls, int(d.get("Instructions Executed", "0") or 0)]
total = sum(v[1] for v in lines.values())
print(f"{first_kernel}: {total} samples over {len(lines)} source lines")
per_file = defaultdict(int)
for (f, _), v in lines.items():
    per_file[f] += v[1]
print("by file:", {f: f"{100 * s / total:.1f}%" for f, s in sorted(per_file.items(), key=lambda kv: -kv[1])})
for (f, ln), (src, s, st, inst) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    reasons = ", ".join(f"{k} {100 * v / s:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * s / total:5.1f}%  {f}:{ln:<4d} inst {inst:>9d}  [{reasons}]  {src}")
