cd $GRAFT_REPO_ROOT
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_n4.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke_n4.log
tail -n 6 gpurun_out/r2_smoke_n4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n4.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_n4.json | head -3
grep -o '"collectives_ms".*' gpurun_out/r2_bench_n4.json | cut -c1-1200
