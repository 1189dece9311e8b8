cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_dist.py -x -q -m gpu > gpurun_out/r2b_dist_tests.log 2>&1; echo "dist tests rc=$?"
tail -n 4 gpurun_out/r2b_dist_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke_n2.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2b_smoke_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2b_bench_n2.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2b_bench_n2.json | head -3
grep -o '"parity_check".*"roofline"' gpurun_out/r2b_bench_n2.json | cut -c1-400
