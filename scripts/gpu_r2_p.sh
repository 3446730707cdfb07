#!/bin/bash
# final profiling of the round-2 tree: ncu launch list of the bench command (+ DRAM bytes), ncu --set full of the conv kernels
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bf16.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_p_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_fwd_tc_kernel|conv_wgrad_tc_kernel" -c 40 -o gpurun_out/r2_conv_bf16 -f python scripts/prof_kernels.py --reps 1 --only conv > gpurun_out/r2_p_ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"window_attn" -c 6 -o gpurun_out/r2_window_attn -f python -m pytest tests/test_gpu_sptr.py -m gpu -q -x -k "end_to_end" > gpurun_out/r2_p_ncu_wa.log 2>&1
ls -la gpurun_out/*.ncu-rep; wc -l gpurun_out/r2_launches_bf16.csv
