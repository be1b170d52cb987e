# round 2, call k: the surface machine after the class split / code-size cut: parity, c1box, and the C2 grids on the C1 film (per-lane cost at 256^3)
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_surfaces.py -q -m gpu -x > gpurun_out/${TAG}_pytest_surf.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_surf.log
(timeout 200 python bench.py --config c1box --steps 10 --warmup 3 --no-e2e 2>gpurun_out/${TAG}_bench_c1box.err | tail -1) > gpurun_out/${TAG}_bench_c1box.json
(timeout 300 python bench.py --config c2box --res 128 --spp 16 --steps 3 --warmup 1 --no-e2e --no-baseline-method --cpu-sample-spp 1 2>gpurun_out/${TAG}_bench_c2grid_c1film.err | tail -1) > gpurun_out/${TAG}_bench_c2grid_c1film.json
python - <<PY
import json
for f in ('c1box', 'c2grid_c1film'):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_%s.json' % f))
        print(f, 'step %.3f ms' % d['ms_per_step'], 'value %.1f' % d['value'], d['phases_ms_rank0'], 'parity', d['parity'] and d['parity']['pass'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
    except Exception as e:
        print(f, 'failed', e)
PY
tail -n 3 gpurun_out/${TAG}_bench_c1box.err gpurun_out/${TAG}_bench_c2grid_c1film.err
