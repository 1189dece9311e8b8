# final: every GPU test at the final sources; ncu of the averaging kernel at the 8-GPU shard size
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/r2i_gpu_tests.log
F="python tools/quick_prof.py -K 125000 -T 200 -A 3 --flags 32 --steps 3"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:average_kernel -s 3 -c 1 -o gpurun_out/prof_r2i_avg125k $F > gpurun_out/r2i_ncu_avg.log 2>&1; echo "ncu rc=$?"
