cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_tile_kernel.py -x -q -m gpu > gpurun_out/r2_tile_tests3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tile_tests3.log
tail -n 8 gpurun_out/r2_tile_tests3.log
rm -f gpurun_out/r2_tile_prof4.log
for K in 1000000 250000 125000; do
  timeout 120 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 1024 --steps 20 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof4.log
done
timeout 120 python tools/quick_prof.py -K 100000 -T 200 -A 2 --flags 1024 --steps 50 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof4.log
timeout 120 python tools/quick_prof.py -K 10000 -T 200 -A 2 --flags 1024 --steps 50 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof4.log
timeout 120 python tools/quick_prof.py -K 10000 -T 200 -A 1 --flags 1024 --steps 50 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof4.log
timeout 120 python tools/quick_prof.py -K 100000 -T 50 -A 2 --flags 1024 --steps 50 2>/dev/null | tail -n 1 >> gpurun_out/r2_tile_prof4.log
cat gpurun_out/r2_tile_prof4.log
