# quarter-tile step kernel: parity (the step-kernel test file with MPPI_STEP_QUARTER=1) and timings
cd $GRAFT_REPO_ROOT
MPPI_STEP_QUARTER=1 timeout 600 python -m pytest tests/test_gpu_step_kernel.py -x -q -m gpu > gpurun_out/r2b_q_tests.log 2>&1; echo "tests rc=$?"
tail -n 5 gpurun_out/r2b_q_tests.log
: > gpurun_out/r2b_q.jsonl
for K in 125000 250000 500000 1000000; do for Q in 0 1; do
  MPPI_STEP_QUARTER=$Q timeout 90 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 128 --steps 30 2>/dev/null | tail -n 1 | sed "s/^{/{\"Q\": $Q, /" >> gpurun_out/r2b_q.jsonl
done; done
cut -c1-200 gpurun_out/r2b_q.jsonl
