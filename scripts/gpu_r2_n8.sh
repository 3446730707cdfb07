#!/bin/bash
# 8-GPU call: bench at N = 8 and N = 4 (DDP + SyncBN over NVLink peer memory), one NCCL-SyncBN run at N = 8 for comparison
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # N port out [extra args / env]
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 10 --warmup 3 --no-cpu-baseline "${@:4}" 2>gpurun_out/r2_n8_$3.err | grep '^{' | tail -1 > gpurun_out/$3.json
  python -c "
import json; d=json.load(open('gpurun_out/$3.json')); print('$3', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2), d['config'].get('syncbn_transport'), d['clocks'])" || tail -5 gpurun_out/r2_n8_$3.err
}
run 8 29541 r2_bench_bf16_8gpu
run 4 29542 r2_bench_bf16_4gpu
U2_SYNCBN_TRANSPORT=nccl run 8 29543 r2_bench_bf16_8gpu_ncclsyncbn
