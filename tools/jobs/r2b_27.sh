cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_dist.py -x -q -m gpu 2>&1 | tail -n 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29627 bench.py --gpus 2 --steps 20 --warmup 5 --no-nccl-leg --no-weak-probe 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=2', d['ms_per_step'], d['parity_check']['ok'], {k:d['collectives_ms']['p2p'][k] for k in ('push_us','wait_slowest_us','merge_us')})"
