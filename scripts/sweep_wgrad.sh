#!/bin/bash
# wgrad pairs-per-CTA sweep: total wgrad ms/step of the bench workload per setting
for mn in 2048 1024 512; do
  for tg in 1184 2368; do
    echo -n "PAIRS_MIN=$mn TARGET=$tg  "
    U2_WGRAD_PAIRS_MIN=$mn U2_WGRAD_TARGET_CTAS=$tg python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],2), round(d['ms_per_step'],2), round(d['roofline']['all_conv']['wgrad']['ms_per_step'],3))"
  done
done
