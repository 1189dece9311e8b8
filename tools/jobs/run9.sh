cd $GRAFT_REPO_ROOT
timeout 1500 python tools/chain_sweep.py > gpurun_out/r2_chain_sweep.jsonl 2> gpurun_out/r2_chain_sweep.err
tail -n 3 gpurun_out/r2_chain_sweep.err
wc -l gpurun_out/r2_chain_sweep.jsonl
