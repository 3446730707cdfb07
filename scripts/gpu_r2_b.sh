#!/bin/bash
# call B: gather microbenchmark, conv diagnostics (traffic switched off piecewise), ncu --set full on the conv kernels
mkdir -p gpurun_out
timeout 300 scripts/microbench/gather_bench 0,4,5,9,12,13,14 128,384 > gpurun_out/r2_gather_microbench.jsonl 2> gpurun_out/r2_gather_microbench.err
timeout 600 python scripts/diag_conv.py --reps 5 --modes 0,1,2,3,16 > gpurun_out/r2_b_diag.log 2>&1
U2_DEBUG_CONV_TIMING=1 timeout 600 python scripts/diag_conv.py --reps 1 --modes 0 > gpurun_out/r2_b_diag_dbg.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_fwd_tc_kernel|conv_wgrad_tc_kernel" -c 12 -o gpurun_out/r2_conv_bf16 python scripts/prof_kernels.py --reps 1 --only conv > gpurun_out/r2_b_ncu.log 2>&1
tail -4 gpurun_out/r2_b_diag.log
