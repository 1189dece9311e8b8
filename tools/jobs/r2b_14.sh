cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipelined.py tests/test_gpu_variants.py -x -q -m gpu 2>&1 | tail -n 3
for P in 0 1; do for cfg in "125000 200 3 32" "250000 200 3 32" "100000 200 2 0" "100000 200 2 512" "10000 200 2 0" "10000 200 2 512"; do
  set -- $cfg
  MPPI_PDL=$P timeout 90 python tools/quick_prof.py -K $1 -T $2 -A $3 --flags $4 --steps 200 2>/dev/null | tail -n 1 | cut -c1-120 | sed "s/^/PDL=$P /"
done; done
