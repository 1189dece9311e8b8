# what the driver does at round end, at the final sources: GPU tests, smoke, bench (both arms)
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/r2h_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2h_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/r2h_bench_ref.json 2>gpurun_out/r2h_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2>gpurun_out/r2h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2h_bench_ref.json','gpurun_out/r2h_bench.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d.get('impl'), d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('gpu_launches'), (d.get('parity_check') or {}).get('ok'))
            if 'other_chains' in d: print({k:round(v['ms_per_step'],4) for k,v in d['other_chains'].items()})
PY
