cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n4.json 2> gpurun_out/r2f_bench_n4.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2f_bench_n4.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2f_bench_n4.json | head -3
grep -o '"parity_check".*"roofline"' gpurun_out/r2f_bench_n4.json | cut -c1-300
grep -o '"kernels".*"clocks"' gpurun_out/r2f_bench_n4.json | cut -c1-500
