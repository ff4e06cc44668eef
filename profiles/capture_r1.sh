#!/bin/bash
# profiles/capture_r1.sh — the two ncu passes behind profiles/r1_*  (run on a B200 under gpurun from the repo root)
#  1. launch list of the default bench command (gpu__time_duration.sum, --clock-control none): per-kernel share of the step
#  2. --set full capture of the three step kernels at step ~200 of the same command (the state bench.py times)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1_launches.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-poly > gpurun_out/r1_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_contact_sphere_pp|k_gather|k_contact_driver" -s 600 -c 3 \
    -o gpurun_out/r1_full -f python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-poly > gpurun_out/r1_full.log 2>&1
ncu -i gpurun_out/r1_full.ncu-rep --page raw --csv > gpurun_out/r1_full.raw.csv 2>/dev/null
tail -2 gpurun_out/r1_full.log | cut -c1-200
