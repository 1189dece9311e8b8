cd $GRAFT_REPO_ROOT
for L in default nr11; do
  if [ $L = nr11 ]; then export MPPI_B200_LIB=$GRAFT_REPO_ROOT/tools/_build/libmppi_nr11.so; fi
  for K in 1000000 500000; do
    timeout 60 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 128 --steps 30 2>/dev/null | tail -n 1 | cut -c1-100 | sed "s/^/$L /"
  done
done
MPPI_B200_LIB=$GRAFT_REPO_ROOT/tools/_build/libmppi_nr11.so timeout 300 python -m pytest tests/test_gpu_step_kernel.py -x -q -m gpu -k "not auto_chain" 2>&1 | tail -n 2
