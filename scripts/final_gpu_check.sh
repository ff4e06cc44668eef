#!/bin/bash
# One bounded GPU call: the -m gpu suite (new driver-motion tests first) and the default bench line of the committed state.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_survey8f_rank3_driver_motion.py tests/test_survey8f_rank4_inner_bonds.py tests/test_gpu_parity.py tests/test_sim_golden.py tests/test_host_operators.py \
    -m gpu -q --tb=short --timeout 120 -p no:cacheprovider > gpurun_out/r1_final_tests.log 2>&1
echo "pytest rc=$?"
tail -4 gpurun_out/r1_final_tests.log
timeout 100 python bench.py --steps 400 --warmup 20 > gpurun_out/r1_final_bench.json 2> gpurun_out/r1_final_bench.err
echo "bench rc=$?"
cut -c1-300 gpurun_out/r1_final_bench.json
