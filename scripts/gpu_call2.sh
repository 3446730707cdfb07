#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/cpu_profile_step.py > gpurun_out/r2_c2_cpu_profile.log 2>&1
U2_DEBUG_CONV_TIMING=1 timeout 600 python scripts/diag_conv.py --reps 1 > gpurun_out/r2_c2_diag_dbg.log 2>&1
grep -c . gpurun_out/r2_c2_diag_dbg.log
