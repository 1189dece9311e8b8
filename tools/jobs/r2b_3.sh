cd $GRAFT_REPO_ROOT
export MPPI_STEP_QUARTER=1
S="python tools/quick_prof.py -K 500000 -T 200 -A 3 --flags 128 --steps 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qstep_kernel -s 3 -c 1 -o gpurun_out/prof_r2b_qstep500k $S > gpurun_out/r2b_ncu_qstep.log 2>&1
tail -3 gpurun_out/r2b_ncu_qstep.log
