#!/bin/bash
# The ncu passes whose summaries go into profiles/ (B200_PROFILING.md recipe), each after the same command exited 0 without ncu.
# usage (on the GPU box, from the repo root): bash tools/final_profile.sh <tag>      e.g. r2g
tag=${1:-r2g}
out=gpurun_out
mkdir -p $out
export PYR_NO_WARMUP=1       # one render per process: launch k of a kernel is wavefront iteration k
set -x
python tools/profile_step.py 8 dragon > $out/${tag}_c2_plain.log 2>&1 || exit 1
python tools/profile_step.py 2 bdpt_cornell_dragon > $out/${tag}_c5_plain.log 2>&1 || exit 1
# launch lists (kernel SHARES of a step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_c2_launches.csv python tools/profile_step.py 8 dragon > $out/${tag}_c2_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_c5_launches.csv python tools/profile_step.py 2 bdpt_cornell_dragon > $out/${tag}_c5_ncu1.log 2>&1
# whole-run DRAM / L2 traffic of every launch of one C2 render, with that render's own ray counts (-> bytes per ray)
PYR_COUNTS_JSON=$out/${tag}_c2_traffic_counts.json ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_c2_traffic.csv python tools/profile_step.py 2 dragon > $out/${tag}_c2_ncu2.log 2>&1
# full-set captures of steady-state launches (iteration 3 of 8 spp: full pool, incoherent rays)
ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_wave_simple|k_bin_keys|k_bin_scatter" -s 12 -c 4 -o $out/${tag}_c2 -f python tools/profile_step.py 8 dragon > $out/${tag}_c2_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_wave_bd" -s 28 -c 7 -o $out/${tag}_c5 -f python tools/profile_step.py 2 bdpt_cornell_dragon > $out/${tag}_c5_ncu3.log 2>&1
# the GPU BVH build (load of config C5's scene)
PYR_BVH_BUILD=gpu ncu --set full --clock-control none --import-source on -k regex:"k_bvh" -s 61 -c 5 -o $out/${tag}_bvh -f python tools/load_timing.py bdpt_cornell_dragon gpu > $out/${tag}_bvh_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_bvh" -c 200 --csv --log-file $out/${tag}_bvh_launches.csv python tools/load_timing.py bdpt_cornell_dragon gpu > $out/${tag}_bvh_ncu1.log 2>&1
ls -la $out/${tag}_*
