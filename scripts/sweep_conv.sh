timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
timeout 100 python scripts/prof_kernels.py --reps 5 --only conv --math bf16 2>&1 | grep -v "wgrad" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['kernel'][:24], d['ms'], d['TFLOP/s'])
"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d[\"dtype\"], d[\"value\"], d[\"ms_per_step\"], d[\"e2e\"][\"value\"], d[\"roofline\"][\"all_conv\"])"
