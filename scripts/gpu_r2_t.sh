#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python scripts/prof_kernels.py --reps 7 --only pv 2>&1 | grep "voxelize" | cut -c1-200
