#!/bin/bash
# 2 GPUs, default bench of the final tree (coordinate prefetch, DDP without the redundant buffer broadcast), step log
mkdir -p gpurun_out
U2_BENCH_STEPLOG=1 timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_n2d_bench.err | grep '^{' | tail -1 > gpurun_out/r2_n2d_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2_n2d_bench.json')); print('2gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d['host_enqueue_ms_per_step'])"
grep -A11 "steplog" gpurun_out/r2_n2d_bench.err | cut -c1-110 | sed -n '24,36p'
