cd $GRAFT_REPO_ROOT
CMD="python tools/quick_prof.py -K 125000 -T 200 -A 3 --flags 1024 --steps 3"
timeout 120 $CMD > gpurun_out/r2_tile_plain11.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 3 -c 1 -o gpurun_out/prof_tile_v3 $CMD > gpurun_out/r2_tile_ncu11.log 2>&1
tail -n 3 gpurun_out/r2_tile_plain11.log gpurun_out/r2_tile_ncu11.log
