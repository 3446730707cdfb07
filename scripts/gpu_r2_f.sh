#!/bin/bash
mkdir -p gpurun_out
U2_DEBUG_CONV_TIMING=2 timeout 300 python scripts/diag_conv.py --reps 1 --modes 0,3 --shapes 1x64x64,1x192x192,8x512x512 > gpurun_out/r2_f_ps_dbg.log 2>&1
grep "ps dbg" gpurun_out/r2_f_ps_dbg.log
