#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/config0.py > gpurun_out/r2_m_config0.log 2>&1; tail -3 gpurun_out/r2_m_config0.log | cut -c1-600
timeout 900 python scripts/layer_sweep.py --reps 5 > gpurun_out/r2_m_layer_sweep.log 2>&1; tail -3 gpurun_out/r2_m_layer_sweep.log | cut -c1-300
timeout 900 python scripts/prof_kernels.py --reps 7 > gpurun_out/r2_m_prof_kernels.log 2>&1; tail -3 gpurun_out/r2_m_prof_kernels.log | cut -c1-300
timeout 600 python scripts/prof_bn.py > gpurun_out/r2_m_prof_bn.log 2>&1; tail -3 gpurun_out/r2_m_prof_bn.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_bench_parity.py -m gpu -q -x -k "prebuilt" 2>&1 | tail -3
