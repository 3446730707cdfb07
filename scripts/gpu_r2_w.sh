#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/prof_attn.py 2>&1 | tee gpurun_out/r2_w_prof_attn.log | tail -14
