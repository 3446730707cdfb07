#!/bin/bash
mkdir -p gpurun_out
for v in 2 3 4; do
  U2_WGRAD_MAX_CTAS=$v U2_BENCH_LAYERS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_s_bench_$v.json 2> gpurun_out/r2_s_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_s_bench_$v.json')); print('max_ctas_w=$v', round(d['value'],2), round(d['ms_per_step'],2), d['roofline']['all_conv']['wgrad'])"
  grep "wgrad" gpurun_out/r2_s_bench_$v.err | grep "64->  64\|128-> 128\|K= 8\|K= 1" | head -8
done
