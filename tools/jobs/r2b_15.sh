# split rollout kernel: parity through the fused-chain tests, then timing vs the fused kernel
cd $GRAFT_REPO_ROOT
MPPI_SPLIT_ROLLOUT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -x -q -m gpu -k "fused or sampled or chain or random" 2>&1 | tail -n 3
for Sp in 0 1; do for K in 30000 60000 125000 166667 250000 500000; do
  MPPI_SPLIT_ROLLOUT=$Sp timeout 90 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 32 --steps 40 2>/dev/null | tail -n 1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('split=$Sp', d['K'], 'graph %.4f rollout %.4f average %.4f'%(d['graph_ms_per_step'], d['rollout_ms'], d['average_ms']))"
done; done
