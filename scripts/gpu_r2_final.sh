#!/bin/bash
# final verification of the tree: full GPU test suite, smoke(), default bench (with cpu_baseline), reference arm (short)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2_final_tests.log; tail -3 gpurun_out/r2_final_tests.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_final_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_final_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'], d['gpu_launches'], d['clocks'])"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; cut -c1-400 gpurun_out/r2_final_ref.json
