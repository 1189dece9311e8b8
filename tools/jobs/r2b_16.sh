# split rollout kernel: first a canary with a short leash, then parity, then timing
cd $GRAFT_REPO_ROOT
export MPPI_SPLIT_ROLLOUT=1
timeout 40 python tools/quick_prof.py -K 30000 -T 200 -A 3 --flags 32 --steps 5 2>&1 | tail -n 1 | cut -c1-150 || { echo CANARY FAILED; exit 1; }
timeout 40 python tools/quick_prof.py -K 166667 -T 50 -A 2 --flags 32 --steps 5 2>&1 | tail -n 1 | cut -c1-150 || { echo CANARY2 FAILED; exit 1; }
timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -x -q -m gpu -k "fused or sampled or chain or random" 2>&1 | tail -n 3
for Sp in 0 1; do for K in 30000 60000 125000 166667 250000 500000; do
  MPPI_SPLIT_ROLLOUT=$Sp timeout 40 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 32 --steps 40 2>/dev/null | tail -n 1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('split=$Sp', d['K'], 'graph %.4f rollout %.4f average %.4f'%(d['graph_ms_per_step'], d['rollout_ms'], d['average_ms']))
except Exception as e: print('split=$Sp $K FAILED')"
done; done
