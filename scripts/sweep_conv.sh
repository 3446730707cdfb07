for mode in 0 1; do
echo "== CPASYNC_MODE_W=$mode"
U2_CPASYNC_MODE_W=$mode timeout 100 python scripts/prof_kernels.py --reps 5 --only conv --math ${1:-bf16} 2>&1 | grep "conv_wgrad" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['kernel'][:24], d['ms'], d['TFLOP/s'])
"
done
