cd $GRAFT_REPO_ROOT
for i in 1 2; do timeout 120 python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 128 --steps 30 2>/dev/null | tail -n 1; done > gpurun_out/r2_step_prof8.log 2>&1
timeout 300 python -m pytest tests/test_gpu_step_kernel.py -x -q -m gpu 2>&1 | tail -n 3 >> gpurun_out/r2_step_prof8.log
cat gpurun_out/r2_step_prof8.log
