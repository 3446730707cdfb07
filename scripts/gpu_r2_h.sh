#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_h_tests.log
tail -3 gpurun_out/r2_h_tests.log
timeout 300 python scripts/diag_conv.py --reps 5 --modes 0,3 > gpurun_out/r2_h_diag_legacy.log 2>&1
U2_CONV_NPW=4 timeout 300 python scripts/diag_conv.py --reps 5 --modes 0,3 > gpurun_out/r2_h_diag_ps.log 2>&1
U2_DEBUG_CONV_TIMING=1 timeout 300 python scripts/diag_conv.py --reps 1 --modes 0 --shapes 1x64x64,1x192x192,8x512x512 > gpurun_out/r2_h_dbg_legacy.log 2>&1
cat gpurun_out/r2_h_diag_legacy.log gpurun_out/r2_h_diag_ps.log; grep "conv dbg" gpurun_out/r2_h_dbg_legacy.log | grep -v occupancy
