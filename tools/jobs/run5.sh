cd $GRAFT_REPO_ROOT
export MPPI_TILE_DEBUG_SKIP=1
CMD="python tools/quick_prof.py -K 125000 -T 200 -A 3 --flags 1024 --steps 3"
$CMD > gpurun_out/r2_tile_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 3 -c 1 -o gpurun_out/prof_tile_gen $CMD > gpurun_out/r2_tile_ncu5.log 2>&1
tail -n 3 gpurun_out/r2_tile_plain5.log gpurun_out/r2_tile_ncu5.log
