#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sptr.py -m gpu -q -x -s -k spformer_model > gpurun_out/r2_q_tests.log 2>&1
grep -n "^rows \|^worst gradient\|passed\|failed" gpurun_out/r2_q_tests.log | cut -c1-1500
