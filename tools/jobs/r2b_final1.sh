# final single-GPU records of round 2: bench (both arms), launch list, ncu --set full of the step kernel, role trace
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2f_bench_ref_n1.json 2> gpurun_out/r2f_bench_ref_n1.err; echo "ref rc=$?"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-configs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches_bench.csv $B > gpurun_out/r2f_ncu_launches.log 2>&1; echo "launch list rc=$?"
S="python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 128 --steps 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/prof_r2f_step $S > gpurun_out/r2f_ncu_step.log 2>&1; echo "ncu step rc=$?"
F="python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 32 --steps 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:average_kernel -s 3 -c 1 -o gpurun_out/prof_r2f_avg $F > gpurun_out/r2f_ncu_avg.log 2>&1; echo "ncu avg rc=$?"
head -c 600 gpurun_out/r2f_bench_n1.json; echo; head -c 400 gpurun_out/r2f_bench_ref_n1.json
