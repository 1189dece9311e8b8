cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_gpu_tests2.log 2>&1; echo "tests rc=$?"
tail -n 6 gpurun_out/r2b_gpu_tests2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/r2b_smoke.log
for i in 1 2; do timeout 90 python tools/quick_prof.py -K 1000000 -T 200 -A 3 --flags 128 --steps 30 2>/dev/null | tail -n 1 | cut -c1-110; done
