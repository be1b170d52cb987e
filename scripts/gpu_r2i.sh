# round 2, call i: surfaces + area light + inherited PRB adjoint on the GPU: parity tests, the C1 / C2 "box" lines
TAG=${1:-r2i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_surfaces.py tests/test_envmap.py -q -m gpu -x > gpurun_out/${TAG}_pytest_surf.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_surf.log
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_surfaces.py --deselect tests/test_envmap.py > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
(timeout 300 python bench.py --config c1box --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench_c1box.err | tail -1) > gpurun_out/${TAG}_bench_c1box.json
(timeout 900 python bench.py --config c2box 2>gpurun_out/${TAG}_bench_c2box.err | tail -1) > gpurun_out/${TAG}_bench_c2box.json
for f in c1box c2box; do python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_bench_$f.json'))
    print('$f', 'step %.3f ms' % d['ms_per_step'], 'value %.0f' % d['value'], 'roofline', {k: d['roofline'][k] for k in ('bound','kernel','achieved','peak','frac')})
    print('  phases', d['phases_ms_rank0']); print('  parity', d['parity']); print('  cpu', d['cpu_baseline']); print('  e2e', d['e2e'] and d['e2e']['ms_per_step'])
except Exception as e:
    print('$f failed', e)
PY
done
tail -5 gpurun_out/${TAG}_bench_c1box.err gpurun_out/${TAG}_bench_c2box.err
