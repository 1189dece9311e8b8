# round-2 profiler evidence: launch list of the bench command, --set full of the two one-kernel steps
cd $GRAFT_REPO_ROOT
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-configs"
timeout 600 $B > gpurun_out/r2_ncu_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
S="python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 128 --steps 3"
timeout 120 $S > gpurun_out/r2_ncu_plain_step.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/prof_r2_step $S > gpurun_out/r2_ncu_step.log 2>&1
Tk="python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 1024 --steps 3"
timeout 120 $Tk > gpurun_out/r2_ncu_plain_tile.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 3 -c 1 -o gpurun_out/prof_r2_tile $Tk > gpurun_out/r2_ncu_tile.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_bench.csv
