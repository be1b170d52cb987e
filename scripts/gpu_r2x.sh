# round 2, call x: sticky collision rounds (few live lanes in a pool: stay in registers, at most RCV_STICKY_MAX steps): parity, C1 and C2 lines
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "volpathsimple_backward or twostates_per_lane" > gpurun_out/${TAG}_pytest_quick.log 2>&1; rc=$?; tail -1 gpurun_out/${TAG}_pytest_quick.log
if [ $rc -ne 0 ]; then echo "quick parity failed or hung (rc=$rc): stopping"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c1.py tests/test_envmap.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
for v in default nosticky; do
  if [ "$v" = "default" ]; then unset RCV_B200_LIB; else export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so; fi
  (timeout 120 python bench.py --config c1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-drt --no-baseline-method 2>&1 | tail -1) > gpurun_out/${TAG}_bench_c1_$v.json
  (timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --supergrid --no-drt --no-baseline-method 2>&1 | tail -1) > gpurun_out/${TAG}_bench_c2_$v.json
  python - <<PY
import json
for f in ('c1', 'c2'):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_%s_$v.json' % f))
        print('$v', f, 'step %.3f ms' % d['ms_per_step'], d['phases_ms_rank0'])
    except Exception as e:
        print('$v', f, 'failed', e)
PY
done
