#!/bin/bash
# 2-GPU call: SyncBN group-path parity test (2 processes, NCCL) + bench variants
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_syncbn.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_n2_syncbn_test.log
tail -5 gpurun_out/r2_n2_syncbn_test.log
bash scripts/gpu_call_n2.sh 2 2>&1 | tee gpurun_out/r2_n2_variants.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/r2_n2_bench.err | grep '^{' | tail -1 > gpurun_out/r2_bench_bf16_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_bf16_2gpu.json')); print('default 2gpu', d['value'], d['ms_per_step'], d['config'].get('syncbn_transport'), d['config'].get('grad_allreduce'))"
