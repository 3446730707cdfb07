#!/bin/bash
mkdir -p gpurun_out
U2_BENCH_HOSTPROF=gpurun_out/r2_n2c_hostprof.txt timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --quick 2>gpurun_out/r2_n2c_bench.err | grep '^{' | tail -1 | cut -c1-200
head -60 gpurun_out/r2_n2c_hostprof.txt | cut -c1-170
