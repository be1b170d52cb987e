# round 2, call aa: chaining consecutive outside-the-medium pieces in the surface machine (RCV_SF_CHAIN = 1 / 2 / 4), c1box
TAG=${1:-r2aa}
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_surfaces.py -q -m gpu -x 2>&1 | tail -1
B="python bench.py --config c1box --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --no-baseline-method"
for v in default chain1 chain4; do
  if [ "$v" = "default" ]; then unset RCV_B200_LIB; else export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so; fi
  (timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_$v.json')); print('$v', round(d['ms_per_step'],2), d['phases_ms_rank0']['twostate_render'], d['phases_ms_rank0']['adjoint'])"
done
