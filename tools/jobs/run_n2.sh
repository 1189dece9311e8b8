# two GPUs: the K-shard tests, smoke's 2-rank leg, bench at N=2 (p2p, with the NCCL leg)
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 1200 python -m pytest tests/test_dist.py -x -q -m gpu > gpurun_out/r2_dist_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_dist_tests.log
tail -n 30 gpurun_out/r2_dist_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_n2.err
head -c 4000 gpurun_out/r2_bench_n2.json
