#!/bin/bash
# wgrad pairs-per-CTA sweep: total wgrad ms/step of the bench workload per setting
for mn in 4096 2048 1024 512 256; do
  for tg in 592 1184; do
    echo "== PAIRS_MIN=$mn TARGET=$tg"
    U2_WGRAD_PAIRS_MIN=$mn U2_WGRAD_TARGET_CTAS=$tg python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['all_conv']['wgrad'])"
  done
done
