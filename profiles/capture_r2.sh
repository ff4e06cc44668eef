#!/bin/bash
# profiles/capture_r2.sh — the ncu passes behind profiles/r2_* and profiles/ncu_traffic.json (run on a B200 under gpurun, repo root).
#  1. launch lists (gpu__time_duration.sum, --clock-control none) of the default bench command and of the polyhedra workload
#  2. --set full captures of the per-step kernels AT THE OPERATING POINTS bench.py times:
#       A = `bench.py --steps 20 --warmup 5`  (the driver's command): step ~15;   B = `--steps 400 --warmup 20`: step ~220
#     (two matching launches per step: k_contact_sphere_pp, k_gather -> -s = 2 x step index)
#  3. --set full captures of the polyhedron kernels (step kernels at step ~15, neighbour-build kernels of the second rebuild) and of
#     k_nbh_sphere
# scripts/ncu_traffic.py turns the raw CSVs into profiles/ncu_traffic.json + profiles/r2_ncu_*.txt
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --no-poly"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_drum.csv $B --steps 20 --warmup 5 > gpurun_out/r2_launches_drum.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_poly.csv $B --workload poly --particles 500000 --steps 20 --warmup 5 > gpurun_out/r2_launches_poly.log 2>&1
full() {  # tag, kernel regex, skip, count, bench args...
  local tag=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/r2_ncu_$tag -f $B "$@" > gpurun_out/r2_ncu_$tag.log 2>&1
  ncu -i gpurun_out/r2_ncu_$tag.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$tag.raw.csv 2>/dev/null
  rm -f gpurun_out/r2_ncu_$tag.ncu-rep
  echo "$tag: $(tail -1 gpurun_out/r2_ncu_$tag.log | cut -c1-120)"
}
full drumA "k_contact_sphere_pp|k_gather" 30 2 --steps 20 --warmup 5
full drumB "k_contact_sphere_pp|k_gather" 440 2 --steps 400 --warmup 20
full polystep "k_detect_poly_pp|k_law_poly_pp|k_gather" 45 3 --workload poly --particles 500000 --steps 20 --warmup 5
full polynbh "k_nbh_poly_cell|k_nbh_poly_replay" 1 2 --workload poly --particles 500000 --steps 6 --warmup 3
full sphnbh "k_nbh_sphere" 2 2 --steps 6 --warmup 3
