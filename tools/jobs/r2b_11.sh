cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/r2b_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2b_bench_n1.err
python - <<'PY'
import json
for l in open('gpurun_out/r2b_bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d[k] for k in ('value','ms_per_step','gpu_launches','vs_baseline') if k in d})
        print('e2e',d.get('e2e')); print('roofline',d.get('roofline')); print('parity',d.get('parity_check',{}).get('ok')); print('configs',json.dumps(d.get('configs'))[:900])
        print('other',json.dumps(d.get('other_chains'))[:700])
PY
