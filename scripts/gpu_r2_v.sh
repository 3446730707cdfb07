#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_spformer.py > gpurun_out/r2_v_spformer.log 2>&1; tail -3 gpurun_out/r2_v_spformer.log | cut -c1-1200
