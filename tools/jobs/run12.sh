cd $GRAFT_REPO_ROOT
timeout 90 python tools/quick_prof.py -K 500000 -T 20 -A 3 --flags 128 --steps 5 2>/dev/null | tail -n 1 > gpurun_out/r2_step_prof12.log || { echo "STEP KERNEL CANARY FAILED" >> gpurun_out/r2_step_prof12.log; cat gpurun_out/r2_step_prof12.log; exit 1; }
for i in 1 2 3; do timeout 90 python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 128 --steps 30 2>/dev/null | tail -n 1 >> gpurun_out/r2_step_prof12.log; done
cat gpurun_out/r2_step_prof12.log
