for cfg in "2 3" "2 2" "3 2"; do
set -- $cfg
echo "== WGRAD_MAX_CTAS=$1 MIN_STAGES=$2"
U2_WGRAD_MAX_CTAS=$1 U2_WGRAD_MIN_STAGES=$2 timeout 100 python scripts/prof_kernels.py --reps 5 --only conv --math bf16 2>&1 | grep "wgrad" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['kernel'][:24], d['ms'], d['TFLOP/s'])
"
done
