#!/bin/bash
# final tree of round 2 (coordinate prefetch, one-launch re-tiling, attention backward v3): full GPU tests, smoke(), default bench
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2_final3_tests.log; tail -3 gpurun_out/r2_final3_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_final3_smoke.log | cut -c1-300
U2_BENCH_LAYERS=1 U2_BENCH_STEPLOG=1 timeout 150 python bench.py > gpurun_out/r2_final3_bench.json 2> gpurun_out/r2_final3_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_final3_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['all_conv'], d['gpu_launches'], d['cpu_baseline'])"
grep -A11 "steplog" gpurun_out/r2_final3_bench.err | cut -c1-110 | sed -n '25,40p'
