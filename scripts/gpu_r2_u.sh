#!/bin/bash
mkdir -p gpurun_out
ls -la oracle/_ref/
timeout 900 python -m pytest tests/test_gpu_sptr_ref.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2_u_tests.log; tail -30 gpurun_out/r2_u_tests.log | cut -c1-300
