cd $GRAFT_REPO_ROOT
which compute-sanitizer
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r2f_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -n 12 gpurun_out/r2f_memcheck.log
