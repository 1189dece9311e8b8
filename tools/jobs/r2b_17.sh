cd $GRAFT_REPO_ROOT
export MPPI_SPLIT_ROLLOUT=1
timeout 40 python tools/quick_prof.py -K 30000 -T 200 -A 3 --flags 32 --steps 5 2>&1 | tail -n 1 | cut -c1-150 || { echo CANARY FAILED; exit 1; }
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -x -q -m gpu -k "fused or sampled or chain or random" 2>&1 | tail -n 2
for K in 30000 60000 125000 166667 250000 500000; do
  timeout 40 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 32 --steps 40 2>/dev/null | tail -n 1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('split=1', d['K'], 'graph %.4f rollout %.4f average %.4f'%(d['graph_ms_per_step'], d['rollout_ms'], d['average_ms']))
except Exception as e: print('split=1 $K FAILED')"
done
