# half-tile step kernel: parity (the step-kernel test file with MPPI_STEP_HALF=1) and timings
cd $GRAFT_REPO_ROOT
MPPI_STEP_HALF=1 timeout 600 python -m pytest tests/test_gpu_step_kernel.py -x -q -m gpu > gpurun_out/r2b_h_tests.log 2>&1; echo "tests rc=$?"
tail -n 5 gpurun_out/r2b_h_tests.log
: > gpurun_out/r2b_h.jsonl
for K in 125000 250000 500000 1000000; do for G in 0 4 8; do
  if [ $G = 0 ]; then export MPPI_STEP_HALF=0; unset MPPI_HSTEP_GROUPS; else export MPPI_STEP_HALF=1 MPPI_HSTEP_GROUPS=$G; fi
  timeout 90 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 128 --steps 30 2>/dev/null | tail -n 1 | sed "s/^{/{\"G\": $G, /" >> gpurun_out/r2b_h.jsonl
done; done
cut -c1-120 gpurun_out/r2b_h.jsonl
