# BASELINE configs[3] (K sweep 1e3..1e7 x 1d/2d/3d, every chain + the auto chain) and configs[4] (closed loop through the C++ driver)
cd $GRAFT_REPO_ROOT
timeout 900 python tools/k_sweep.py > gpurun_out/r2f_k_sweep.jsonl 2> gpurun_out/r2f_k_sweep.err; echo "sweep rc=$?"
wc -l gpurun_out/r2f_k_sweep.jsonl
for PU in 0 50; do
  timeout 120 cpp/mppi_main -c config/point_mass2d.yaml --samples 100000 --horizon 200 --steps 1000 --plant ideal --plant-us $PU --quiet 2>&1 | tail -n 6 | sed "s/^/plant_us=$PU /"
done
