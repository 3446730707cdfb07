#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_parity.py -m gpu -q -x -k "prefetch or prebuilt" 2>&1 | tail -40 | tee gpurun_out/r2_z_tests.log | tail -5
U2_BENCH_STEPLOG=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_z_bench.err | tee gpurun_out/r2_z_bench.json | cut -c1-200
grep -A11 "steplog" gpurun_out/r2_z_bench.err | tail -80
