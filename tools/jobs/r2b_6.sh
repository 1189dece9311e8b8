# step kernel with the deterministic first round: parity, then the chains at the shard sizes of 2/4/8 GPUs
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_step_kernel.py -x -q -m gpu > gpurun_out/r2b_s_tests.log 2>&1; echo "tests rc=$?"
tail -n 3 gpurun_out/r2b_s_tests.log
timeout 600 python tools/chain_sweep.py --shapes 3:200 --K 125000,166667,250000,333334,500000,1000000 --chains unfused,fused,step > gpurun_out/r2b_sweep.jsonl 2>gpurun_out/r2b_sweep.err
cut -c1-260 gpurun_out/r2b_sweep.jsonl
