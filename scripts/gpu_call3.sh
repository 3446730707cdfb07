#!/bin/bash
mkdir -p gpurun_out
U2_DEBUG_CONV_TIMING=1 timeout 600 python scripts/diag_conv.py --reps 1 --modes 0,8,16,24,48 --shapes 1x64x64,1x192x192,8x512x512 > gpurun_out/r2_c3_diag_dbg.log 2>&1
timeout 600 python scripts/diag_conv.py --reps 5 --modes 0,8,10,16,24 > gpurun_out/r2_c3_diag.log 2>&1
cat gpurun_out/r2_c3_diag.log
