cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_step_kernel.py tests/test_gpu_tile_kernel.py tests/test_gpu_hardening.py -x -q -m gpu > gpurun_out/r2b_merge_tests.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/r2b_merge_tests.log
for K in 1000000 500000 250000; do for i in 1 2; do timeout 90 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 128 --steps 30 2>/dev/null | tail -n 1 | cut -c1-110; done; done
timeout 90 python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 1024 --steps 10 2>/dev/null | tail -n 1 | cut -c1-110
