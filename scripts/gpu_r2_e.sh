#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_e_tests.log
tail -5 gpurun_out/r2_e_tests.log
timeout 300 python scripts/diag_conv.py --reps 5 --modes 0,1,2,3 > gpurun_out/r2_e_diag_ps.log 2>&1
U2_CONV_KERNEL=legacy timeout 300 python scripts/diag_conv.py --reps 5 --modes 0 > gpurun_out/r2_e_diag_legacy.log 2>&1
cat gpurun_out/r2_e_diag_ps.log gpurun_out/r2_e_diag_legacy.log
