# pipe throughput microbenchmark + small-shard kernel times + ncu of the fused rollout at 125k
cd $GRAFT_REPO_ROOT
timeout 120 tools/_build/pipes > gpurun_out/r2b_pipes.txt 2>&1; echo "pipes rc=$?"
: > gpurun_out/r2b_small.jsonl
for K in 10000 60000 125000 250000; do for F in 0 32 128; do
  timeout 90 python tools/quick_prof.py -K $K -T 200 -A 3 --flags $F --steps 30 2>/dev/null | tail -n 1 >> gpurun_out/r2b_small.jsonl
done; done
S="python tools/quick_prof.py -K 125000 -T 200 -A 3 --flags 32 --steps 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o gpurun_out/prof_r2b_fused125k $S > gpurun_out/r2b_ncu_fused.log 2>&1
cat gpurun_out/r2b_pipes.txt; cat gpurun_out/r2b_small.jsonl
