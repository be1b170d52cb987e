# round 2, call h: envmap emitter on the GPU (parity), full GPU suite, default bench (regression check of the headline)
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
(timeout 600 python bench.py --no-cpu-baseline --supergrid 2>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_c2.json
python - <<PY
import json
d = json.load(open('gpurun_out/${TAG}_bench_c2.json'))
print('c2 step %.3f ms' % d['ms_per_step'], d['phases_ms_rank0'], 'drt', d['drt'] and d['drt']['ms_per_step'])
PY
tail -3 gpurun_out/${TAG}_bench.err
