#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_o_tests.log
tail -3 gpurun_out/r2_o_tests.log
U2_BENCH_LAYERS=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_o_bench.json 2> gpurun_out/r2_o_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2_o_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['host_enqueue_ms_per_step'], d['gpu_launches'], d['roofline']['frac'])"
grep "4->  64" gpurun_out/r2_o_bench.err
