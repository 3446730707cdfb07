#!/bin/bash
# round-2 re-entry call A: GPU tests, default bench, ncu launch list of the same command
mkdir -p gpurun_out
{ python -c "import torchsparse; print('torchsparse importable', getattr(torchsparse,'__version__','?'))"; ls -la baseline/_ref; nvidia-smi -L; nproc; } > gpurun_out/r2_probe.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s -x 2>&1 | tail -200 > gpurun_out/r2_a_tests.log
U2_BENCH_LAYERS=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_a_bench.json 2> gpurun_out/r2_a_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_a_ncu_bench.log 2>&1
tail -5 gpurun_out/r2_a_tests.log; tail -c 3000 gpurun_out/r2_a_bench.json
