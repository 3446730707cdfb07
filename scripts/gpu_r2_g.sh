#!/bin/bash
mkdir -p gpurun_out
timeout 300 scripts/microbench/gather_bench 0,10,5 128 > gpurun_out/r2_gather_microbench_hybrid.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_gather_microbench_hybrid.jsonl'):
    try: r=json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['pitch'], r['zero_fill'], 'm',r['method'],'W',r['W'],'ctas',r['ctas'],'slotB/clk',r['slot_B_per_clk_sm'],'cyc/item',r['cycles_per_item_sm'])
PY
