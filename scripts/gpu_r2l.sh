# round 2, call l: how the surface machine scales with the number of lanes (c2box grids; film / spp grow)
TAG=${1:-r2l}
mkdir -p gpurun_out
for cfg in "256 16" "256 64" "512 16"; do
  set -- $cfg
  (timeout 150 python bench.py --config c2box --res $1 --spp $2 --steps 2 --warmup 1 --no-e2e --no-baseline-method --no-cpu-baseline 2>gpurun_out/${TAG}_res$1_spp$2.err | tail -1) > gpurun_out/${TAG}_res$1_spp$2.json
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_res$1_spp$2.json'))
    print('res $1 spp $2', 'step %.1f ms' % d['ms_per_step'], 'value %.1f' % d['value'], d['phases_ms_rank0'])
except Exception as e:
    print('res $1 spp $2 failed', e)
PY
  tail -n 2 gpurun_out/${TAG}_res$1_spp$2.err
done
