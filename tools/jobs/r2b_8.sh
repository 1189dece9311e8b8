# average kernel: chunk-major (newest rows first) vs slab-major at small shards; fused store hints
cd $GRAFT_REPO_ROOT
: > gpurun_out/r2b_avg.jsonl
run() { # tag, env...
  tag=$1; shift
  for cfg in "125000 200 3 32" "250000 200 3 32" "125000 200 3 0" "100000 200 2 0" "100000 200 2 32" "50000 200 2 0" "10000 200 2 0"; do
    set -- $cfg
    env $ENVV timeout 90 python tools/quick_prof.py -K $1 -T $2 -A $3 --flags $4 --steps 40 2>/dev/null | tail -n 1 | sed "s/^{/{\"tag\": \"$tag\", /" >> gpurun_out/r2b_avg.jsonl
  done
}
ENVV="MPPI_AVG_CHUNK_MAJOR=0" run slab
ENVV="MPPI_AVG_CHUNK_MAJOR=1" run chunk
ENVV="MPPI_AVG_CHUNK_MAJOR=1 MPPI_B200_LIB=$GRAFT_REPO_ROOT/tools/_build/libmppi_wt.so" run chunk_wt
ENVV="MPPI_AVG_CHUNK_MAJOR=1 MPPI_B200_LIB=$GRAFT_REPO_ROOT/tools/_build/libmppi_cs.so" run chunk_cs
ENVV="MPPI_AVG_CHUNK_MAJOR=0 MPPI_B200_LIB=$GRAFT_REPO_ROOT/tools/_build/libmppi_wt.so" run slab_wt
python - <<'PY'
import json
for l in open('gpurun_out/r2b_avg.jsonl'):
    d=json.loads(l)
    print(d['tag'],d['K'],d['A'],d['flags'],'graph %.4f'%d['graph_ms_per_step'],{k:round(d[k],4) for k in ('sample_ms','rollout_ms','average_ms') if k in d})
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipelined.py -x -q -m gpu 2>&1 | tail -n 3
