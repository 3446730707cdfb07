#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/cpu_profile_step.py > gpurun_out/r2_j_cpu_profile.log 2>&1
head -60 gpurun_out/r2_j_cpu_profile.log
