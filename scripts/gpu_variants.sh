mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2_pytest.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
(timeout 300 $B 2>&1 | tail -1) > gpurun_out/r2_default.json
for v in b256m3 b128m5 b256m4 b128m6 th8 th16 th28 th32; do
  (RCV_B200_LIB=$PWD/build/variants/librcv_$v.so timeout 300 $B 2>&1 | tail -1) > gpurun_out/r2_$v.json
done
(timeout 300 $B --tracer 1 2>&1 | tail -1) > gpurun_out/r2_simple.json
(timeout 300 $B --tracer 2 2>&1 | tail -1) > gpurun_out/r2_batch.json
tail -3 gpurun_out/r2_pytest.log
