cd $GRAFT_REPO_ROOT
for L in libmppi_b200 skip8 skip4; do
  if [ $L != libmppi_b200 ]; then export MPPI_B200_LIB=$GRAFT_REPO_ROOT/tools/_build/libmppi_$L.so; fi
  for K in 1000000 500000; do
    timeout 90 python tools/quick_prof.py -K $K -T 200 -A 3 --flags 128 --steps 30 2>/dev/null | tail -n 1 | cut -c1-100 | sed "s/^/$L /"
  done
done
