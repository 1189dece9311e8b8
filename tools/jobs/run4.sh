cd $GRAFT_REPO_ROOT
for K in 1000000 125000; do
  MPPI_TILE_DEBUG_SKIP=1 timeout 120 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 1024 --steps 20 2>/dev/null | tail -1
done > gpurun_out/r2_tile_prof3.log 2>&1
cat gpurun_out/r2_tile_prof3.log
