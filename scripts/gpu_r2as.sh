# round 2, call as: the adjoint's film-weight pass on a side stream during the two-state render, also at N = 1 (bench step)
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --supergrid --no-baseline-method"
for k in 1 2; do
  (timeout 100 $B 2>&1 | tail -1) > gpurun_out/r2as_quick$k.json
  python -c "
import json
d=json.load(open('gpurun_out/r2as_quick$k.json')); print('quick$k', round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['phases_ms_rank0'].items()})"
done
(timeout 600 python bench.py 2>gpurun_out/r2as_bench.err | tail -1) > gpurun_out/r2as_bench_n1.json
python -c "
import json
d=json.load(open('gpurun_out/r2as_bench_n1.json'))
print('ours', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e_device_optimizer']['ms_per_step'], 'launches', d['gpu_launches'], 'parity', d['parity']['pass'], d['roofline']['frac'], d['clocks'], d['phases_ms_rank0'])"
tail -n 2 gpurun_out/r2as_bench.err
