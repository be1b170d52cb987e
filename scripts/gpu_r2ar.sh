# round 2, call ar: racecheck of smoke(); memcheck of the surface / envmap / exact-C1 GPU tests
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ar_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2ar_racecheck.log | cut -c1-200
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_surfaces.py tests/test_envmap.py tests/test_gpu_c1.py -m gpu -q -x > gpurun_out/r2ar_memcheck_more.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2ar_memcheck_more.log | cut -c1-200
