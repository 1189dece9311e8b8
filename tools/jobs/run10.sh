cd $GRAFT_REPO_ROOT
rm -f gpurun_out/r2_tile_prof5.log
# canaries first: a kernel that deadlocks must cost one minute, not ten
timeout 90 python tools/quick_prof.py -K 500000 -T 20 -A 3 --flags 128 --steps 5 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof5.log || { echo "STEP KERNEL CANARY FAILED" >> gpurun_out/r2_tile_prof5.log; cat gpurun_out/r2_tile_prof5.log; exit 1; }
timeout 90 python tools/quick_prof.py -K 20000 -T 30 -A 3 --flags 1024 --steps 5 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof5.log || { echo "TILE KERNEL CANARY FAILED" >> gpurun_out/r2_tile_prof5.log; cat gpurun_out/r2_tile_prof5.log; exit 1; }
timeout 400 python -m pytest tests/test_gpu_tile_kernel.py -x -q -m gpu --timeout 120 > gpurun_out/r2_tile_tests4.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tile_tests4.log
tail -n 8 gpurun_out/r2_tile_tests4.log
for K in 1000000 250000 125000; do
  timeout 90 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 1024 --steps 20 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof5.log
done
timeout 90 python tools/quick_prof.py -K 100000 -T 200 -A 2 --flags 1024 --steps 50 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof5.log
timeout 90 python tools/quick_prof.py -K 10000 -T 200 -A 2 --flags 1024 --steps 50 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof5.log
timeout 90 python tools/quick_prof.py -K 10000 -T 200 -A 1 --flags 1024 --steps 50 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof5.log
timeout 90 python tools/quick_prof.py -K 100000 -T 50 -A 2 --flags 1024 --steps 50 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof5.log
for i in 1 2; do timeout 90 python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 128 --steps 30 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof5.log; done
timeout 300 python -m pytest tests/test_gpu_step_kernel.py -x -q -m gpu --timeout 120 2>&1 | tail -n 3 >> gpurun_out/r2_tile_prof5.log
cat gpurun_out/r2_tile_prof5.log
