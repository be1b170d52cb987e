# round 2, call aq: compute-sanitizer memcheck over the GPU parity tests (as many as fit in two minutes)
mkdir -p gpurun_out
timeout 130 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2aq_memcheck_parity.log 2>&1; echo "memcheck rc=$?"
grep -c PASSED gpurun_out/r2aq_memcheck_parity.log; tail -6 gpurun_out/r2aq_memcheck_parity.log | cut -c1-200
