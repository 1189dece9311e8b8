set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_tile_kernel.py -x -q -m gpu > gpurun_out/r2_tile_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tile_tests.log
tail -30 gpurun_out/r2_tile_tests.log
for K in 1000000 250000 125000; do
 for F in 1024 256; do
  timeout 120 python tools/quick_prof.py -K $K -T 200 -A 3 --flags $F --steps 20 2>/dev/null | tail -1
 done
done >> gpurun_out/r2_tile_prof.log 2>&1
for F in 1024 256; do timeout 120 python tools/quick_prof.py -K 100000 -T 200 -A 2 --flags $F --steps 50 2>/dev/null | tail -1; done >> gpurun_out/r2_tile_prof.log 2>&1
for F in 1024 256; do timeout 120 python tools/quick_prof.py -K 10000 -T 200 -A 2 --flags $F --steps 50 2>/dev/null | tail -1; done >> gpurun_out/r2_tile_prof.log 2>&1
cat gpurun_out/r2_tile_prof.log
