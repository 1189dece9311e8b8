# full single-GPU validation: every GPU test, smoke, bench N=1 (both arms)
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_gpu_tests.log
tail -n 6 gpurun_out/r2_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
tail -n 8 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n1.err
head -c 3000 gpurun_out/r2_bench_n1.json
