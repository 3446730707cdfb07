#!/bin/bash
mkdir -p gpurun_out
run() { # N port tag args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 10 --warmup 3 --no-cpu-baseline "${@:4}" 2>gpurun_out/r2_n8b_$3.err | grep '^{' | tail -1 > gpurun_out/r2_n8b_$3.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_n8b_$3.json')); print('$3', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2), d['config'].get('syncbn_transport'), d['config'].get('grad_allreduce'))" || tail -5 gpurun_out/r2_n8b_$3.err
}
run 8 29551 bf16grad --grad-bf16
run 8 29552 bf16grad_bucket100 --grad-bf16 --bucket-mb 100
run 8 29553 nosyncbn --no-sync-bn
run 8 29554 bucket100 --bucket-mb 100
