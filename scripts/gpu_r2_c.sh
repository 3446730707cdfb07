#!/bin/bash
mkdir -p gpurun_out
U2_DEBUG_CONV_TIMING=1 timeout 600 python scripts/diag_conv.py --reps 1 --modes 0,3 --shapes 1x64x64,1x192x192,8x512x512 > gpurun_out/r2_c_diag_dbg.log 2>&1
U2_CONV_MAX_CTAS=1 timeout 600 python scripts/diag_conv.py --reps 5 --modes 0,3 --shapes 1x64x64,1x192x192,8x512x512 > gpurun_out/r2_c_diag_1cta.log 2>&1
grep "conv dbg\]" gpurun_out/r2_c_diag_dbg.log | grep -v occupancy
cat gpurun_out/r2_c_diag_1cta.log
