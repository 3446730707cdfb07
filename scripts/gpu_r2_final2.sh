#!/bin/bash
# final tree: full GPU tests, smoke(), per-kernel rooflines, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2_final_tests.log; tail -3 gpurun_out/r2_final_tests.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r2_final_smoke.log | cut -c1-300
timeout 900 python scripts/prof_kernels.py --reps 7 > gpurun_out/r2_m_prof_kernels.log 2>&1
timeout 600 python scripts/prof_bn.py > gpurun_out/r2_m_prof_bn.log 2>&1
U2_BENCH_LAYERS=1 timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_final_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['all_conv'], d['gpu_launches'])"
