cd $GRAFT_REPO_ROOT
: > gpurun_out/r2b_p7.jsonl
for L in libmppi_b200 p7; do
  if [ $L = p7 ]; then export MPPI_B200_LIB=$GRAFT_REPO_ROOT/tools/_build/libmppi_p7.so; fi
  for cfg in "1000000 128" "1000000 32" "1000000 0" "125000 32" "250000 32" "500000 128"; do
    set -- $cfg
    timeout 90 python tools/quick_prof.py -K $1 -T 200 -A 3 --flags $2 --steps 30 2>/dev/null | tail -n 1 | sed "s/^{/{\"lib\": \"$L\", /" >> gpurun_out/r2b_p7.jsonl
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2b_p7.jsonl'):
    d=json.loads(l)
    print(d['lib'],d['K'],d['flags'],'graph %.4f'%d['graph_ms_per_step'],{k:round(d[k],4) for k in ('sample_ms','rollout_ms','average_ms') if k in d})
PY
