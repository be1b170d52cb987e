# round 2, call an: suspended supergrid walks in the two-state pool kernel (RCV_WALK_CAP 0 = none / 6 / 12 / 20)
TAG=${1:-r2an}
mkdir -p gpurun_out
export RCV_B200_LIB=$PWD/build/variants/librcv_walk6.so
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --supergrid --no-baseline-method --majorant-factor 8"
for v in walk0 walk6 walk12 walk20; do
  export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so
  (timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_c2sg8_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_c2sg8_$v.json')); print('c2 sg8 $v', round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['phases_ms_rank0'].items()}, d['parity']['status'] if d.get('parity') else None)"
done
B="python bench.py --config refdefault --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-baseline-method"
for v in walk0 walk6 walk12 walk20; do
  export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so
  (timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_refdefault_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_refdefault_$v.json')); print('refdefault $v', round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['phases_ms_rank0'].items()}, d['parity']['status'] if d.get('parity') else None)"
done
