#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2_r_tests.log; tail -12 gpurun_out/r2_r_tests.log | cut -c1-250
U2_BENCH_LAYERS=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_r_bench.json 2> gpurun_out/r2_r_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2_r_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['all_conv'])"
grep wgrad gpurun_out/r2_r_bench.err | head -14
