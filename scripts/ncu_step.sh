#!/bin/bash
# ncu full capture of the per-step kernels (one launch each, after warm-up): bash ncu_step.sh <tag> [extra bench args]
tag=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_contact_sphere_pp|k_gather" -s 20 -c 2 \
  -o gpurun_out/ncu_$tag -f python bench.py --steps 12 --warmup 10 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_$tag.raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_$tag.log
