# round 2, call ao: round-end validation of the in-tree build (as gpu_final.sh) plus the other named lines
bash scripts/gpu_final.sh r2ao
for c in c1 c1box refdefault; do
  (timeout 200 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2ao_$c.err | tail -1) > gpurun_out/r2ao_bench_$c.json
  python -c "
import json
d=json.load(open('gpurun_out/r2ao_bench_$c.json')); print('$c', round(d['ms_per_step'],2), d['value'], {k: round(x,2) for k,x in d['phases_ms_rank0'].items()}, 'parity', (d.get('parity') or {}).get('pass'))"
done
