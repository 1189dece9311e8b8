# chain sweep over shard sizes (incl. the 3/5/6/7-GPU shards of 1e6) after the CTA-size rule and the cost-model policy
cd $GRAFT_REPO_ROOT
timeout 1500 python tools/chain_sweep.py --shapes 3:200,2:200,1:200,2:50,4:100,2:120,2:10 --K 3000,30000,50000,100000,125000,142858,150000,166667,200000,250000,333334,400000,450000,500000,700000,1000000 --chains unfused,fused,step > gpurun_out/r2b_sweep2.jsonl 2>gpurun_out/r2b_sweep2.err
python - <<'PY'
import json
for l in open('gpurun_out/r2b_sweep2.jsonl'):
    d=json.loads(l)
    print(d['A'],d['T'],d['K'],{k:round(v,4) for k,v in d['ms'].items()},d['best'],d['auto'],round(d['auto_vs_best'],3))
PY
