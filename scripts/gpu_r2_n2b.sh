#!/bin/bash
# 2-GPU call: prefetch test again + bench with and without the coordinate prefetch (DDP + peer-memory SyncBN)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_parity.py -m gpu -q -x -k "prefetch or prebuilt" 2>&1 | tail -30 | tee gpurun_out/r2_n2b_tests.log | tail -3
for v in "" "--no-prefetch"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline $v 2>gpurun_out/r2_n2b_bench.err | grep '^{' | tail -1 > gpurun_out/r2_n2b_bench$v.json
python -c "
import json; d=json.load(open('gpurun_out/r2_n2b_bench$v.json')); print('2gpu $v', d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('syncbn_transport'), d['config'].get('coord_prefetch'))"
done
