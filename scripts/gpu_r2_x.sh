#!/bin/bash
# attention backward v2: parity (oracle + the reference's own kernels), kernel timing, model step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sptr.py tests/test_gpu_sptr_ref.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2_x_tests.log
timeout 600 python scripts/prof_attn.py 2>&1 | tee gpurun_out/r2_x_prof_attn.log | tail -12
timeout 900 python scripts/bench_spformer.py 2>&1 | tee gpurun_out/r2_x_bench_spformer.log | tail -6
