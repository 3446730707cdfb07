#!/bin/bash
# N-GPU variants of the bench: SyncBN transport, gradient all-reduce dtype, bucket size (N = $1, default 2)
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "${@:3}" 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', round(d['value'],2), round(d['ms_per_step'],2), d['config'].get('syncbn_transport'), d['config'].get('grad_allreduce'))"; }
U2_SYNCBN_TRANSPORT=nccl run 29521 nccl-syncbn
run 29522 peer-syncbn
run 29523 peer-syncbn+bf16grad --grad-bf16
run 29524 peer-syncbn+bf16grad+bucket100 --grad-bf16 --bucket-mb 100
run 29526 peer-syncbn+bucket100 --bucket-mb 100
run 29525 no-syncbn --no-sync-bn
