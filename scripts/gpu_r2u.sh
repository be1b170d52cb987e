# round 2, call u: 256-bit stores in the brick builder (rotation phase), and the C4 shape on one GPU with its event counts
TAG=${1:-r2u}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugins.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
(timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --supergrid --no-baseline-method 2>&1 | tail -1) > gpurun_out/${TAG}_bench_c2.json
(timeout 600 python bench.py --config c4 --spp 64 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-drt --no-baseline-method 2>gpurun_out/${TAG}_c4.err | tail -1) > gpurun_out/${TAG}_bench_c4_64spp.json
python - <<PY
import json
for f in ('c2', 'c4_64spp'):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_%s.json' % f))
        print(f, 'step %.3f ms' % d['ms_per_step'], 'value %.0f' % d['value'], d['phases_ms_rank0'])
        for k, v in d['kernels'].items(): print('   ', k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
    except Exception as e:
        print(f, 'failed', e)
PY
tail -n 3 gpurun_out/${TAG}_c4.err
