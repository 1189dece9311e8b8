# DRAM traffic of the three dominant kernels at the final sources (profiles/*_traffic.json)
cd $GRAFT_REPO_ROOT
S="python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 128 --steps 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/prof_r2g_step $S > gpurun_out/r2g_ncu_step.log 2>&1; echo "ncu step rc=$?"
F="python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 32 --steps 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:average_kernel -s 3 -c 1 -o gpurun_out/prof_r2g_avg $F > gpurun_out/r2g_ncu_avg.log 2>&1; echo "ncu avg rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o gpurun_out/prof_r2g_fused $F > gpurun_out/r2g_ncu_fused.log 2>&1; echo "ncu fused rc=$?"
Tk="python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 1024 --steps 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 3 -c 1 -o gpurun_out/prof_r2g_tile $Tk > gpurun_out/r2g_ncu_tile.log 2>&1; echo "ncu tile rc=$?"
