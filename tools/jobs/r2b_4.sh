cd $GRAFT_REPO_ROOT
for K in 125000 250000; do timeout 120 python tools/step_trace.py $K > gpurun_out/r2b_trace_$K.txt 2>&1; done
grep -E "all CTAs|merge|finalize" gpurun_out/r2b_trace_125000.txt gpurun_out/r2b_trace_250000.txt
