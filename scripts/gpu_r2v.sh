# round 2, call v: the L2-resident grids (64^3) with enough lanes that the launch is throughput-bound, not tail-bound
TAG=${1:-r2v}
mkdir -p gpurun_out
(timeout 300 python bench.py --config c1 --res 512 --spp 64 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --no-baseline-method 2>gpurun_out/${TAG}.err | tail -1) > gpurun_out/${TAG}_bench_c1grid_c2film.json
python - <<PY
import json
d = json.load(open('gpurun_out/${TAG}_bench_c1grid_c2film.json'))
print('step %.3f ms' % d['ms_per_step'], d['phases_ms_rank0'], d['roofline'])
for k, v in d['kernels'].items(): print('   ', k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
PY
tail -n 3 gpurun_out/${TAG}.err
