#!/bin/bash
# A/B of the conv gather path: LDGSTS (.ca) vs register-staged LDG+STS (U2_CPASYNC_MODE=9)
for mode in 1 9; do
  echo "== U2_CPASYNC_MODE=$mode"
  U2_CPASYNC_MODE=$mode timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "bf16 or tf32" 2>&1 | tail -1
  U2_CPASYNC_MODE=$mode timeout 300 python scripts/prof_kernels.py --only conv --reps 5 2>&1 | grep "conv_fwd\|conv_dgrad" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['kernel'][:40], d['ms'], d.get('TFLOP/s'))"
done
