# round 2, call ap: compute-sanitizer (memcheck, synccheck) over smoke() — small volume + boxed scene + envmap through the in-tree build
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ap_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2ap_memcheck.log
timeout 200 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ap_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/r2ap_synccheck.log
