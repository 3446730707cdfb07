#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sptr.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2_n_tests.log
tail -40 gpurun_out/r2_n_tests.log
