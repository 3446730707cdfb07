timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
for tm in 1 4; do
echo "== WGRAD_TM=$tm"
U2_WGRAD_TM=$tm timeout 100 python scripts/prof_kernels.py --reps 5 --only conv --math ${1:-bf16} 2>&1 | grep "conv_wgrad" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['kernel'][:24], d['ms'], d['TFLOP/s'])
"
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"dtype\"], d[\"value\"], d[\"ms_per_step\"], d[\"e2e\"][\"value\"], d[\"roofline\"][\"all_conv\"])"
