# round 2, call t: warp-aggregated film splats / film weights: full GPU suite, reference-default and default lines
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
(timeout 400 python bench.py --config refdefault --no-cpu-baseline --no-e2e 2>gpurun_out/${TAG}_refdefault.err | tail -1) > gpurun_out/${TAG}_bench_refdefault.json
(timeout 400 python bench.py --no-cpu-baseline --supergrid 2>gpurun_out/${TAG}_c2.err | tail -1) > gpurun_out/${TAG}_bench_c2.json
python - <<PY
import json
for f in ('refdefault', 'c2'):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_%s.json' % f))
        print(f, 'step %.3f ms' % d['ms_per_step'], d['phases_ms_rank0'], 'e2e_dev', d['e2e_device_optimizer'] and d['e2e_device_optimizer']['ms_per_step'])
    except Exception as e:
        print(f, 'failed', e)
PY
tail -n 3 gpurun_out/${TAG}_refdefault.err gpurun_out/${TAG}_c2.err
