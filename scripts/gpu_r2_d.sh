#!/bin/bash
mkdir -p gpurun_out
U2_BENCH_LAYERS=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_d_bench.json 2> gpurun_out/r2_d_bench.err
cat gpurun_out/r2_d_bench.err | head -70
