export U2_NO_TMA_GATHER=1
timeout 200 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
for npw in 4 8 16; do for tmax in 1 2 4; do
echo "== NPW=$npw TMAX=$tmax"
U2_CONV_NPW=$npw U2_CONV_TMAX=$tmax timeout 100 python scripts/prof_kernels.py --reps 5 --only conv 2>&1 | grep "conv_fwd" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d['kernel'][:24], d['ms'], d['TFLOP/s'])
"
done; done
