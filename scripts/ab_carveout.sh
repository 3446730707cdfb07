#!/bin/bash
# does asking for the full shared-memory carve-out change how many conv CTAs are co-resident?
U2_DEBUG_CONV_TIMING=1 python scripts/prof_kernels.py --only conv --reps 1 2>&1 | grep "conv dbg" | sort | uniq -c | sort -rn | head -12
for v in 1 ""; do
  echo "== U2_NO_CARVEOUT=$v"
  U2_NO_CARVEOUT=$v python scripts/prof_kernels.py --only conv --reps 5 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['kernel'][:44], d['ms'], d.get('TFLOP/s'))"
done
