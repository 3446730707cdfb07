#!/bin/bash
# one-launch weight re-tiling: tests, bench, host profile of a step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bench_parity.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_y_tests.log
U2_BENCH_HOSTPROF=gpurun_out/r2_y_hostprof.txt timeout 900 python bench.py --steps 6 --warmup 3 2>gpurun_out/r2_y_bench.err | tee gpurun_out/r2_y_bench.json | cut -c1-900
