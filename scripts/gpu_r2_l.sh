#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bench_parity.py -m gpu -q -x -k "bf16x3 or prebuilt or persistent" 2>&1 | tail -30 > gpurun_out/r2_l_tests.log
tail -12 gpurun_out/r2_l_tests.log
for m in bf16x3 tf32 fp32; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --math $m > gpurun_out/r2_l_bench_$m.json 2> gpurun_out/r2_l_bench_$m.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_l_bench_$m.json')); print('$m', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['all_conv'])"
done
