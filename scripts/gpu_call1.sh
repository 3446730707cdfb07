#!/bin/bash
# round-2 call 1: probe, tests, conv traffic diagnostics, step profile, bench
mkdir -p gpurun_out
{ python -c "import torchsparse; print('torchsparse importable', getattr(torchsparse,'__version__','?'))"; ls -la baseline/_ref; nvidia-smi -L; nproc; } > gpurun_out/r2_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2_c1_tests.log
timeout 600 python scripts/diag_conv.py --reps 5 > gpurun_out/r2_c1_diag.log 2>&1
timeout 600 python scripts/profile_step.py > gpurun_out/r2_c1_profile_step.log 2>&1
U2_BENCH_LAYERS=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err
tail -5 gpurun_out/r2_c1_tests.log; cat gpurun_out/r2_c1_diag.log | tail -12
