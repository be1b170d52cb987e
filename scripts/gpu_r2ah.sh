# round 2, call ah: the supergrid walk with a linear cell index, steps-left counters and select-based stepping (ddalin) vs the (cx, cy, cz) loop (dda1)
TAG=${1:-r2ah}
mkdir -p gpurun_out
export RCV_B200_LIB=$PWD/build/variants/librcv_ddalin.so
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --supergrid --no-baseline-method --majorant-factor 8"
for v in dda1 ddalin dda1 ddalin; do
  export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so
  (timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_c2sg8_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_c2sg8_$v.json')); print('c2 sg8 $v', round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['phases_ms_rank0'].items()}, d['parity']['status'] if d.get('parity') else None)"
done
B="python bench.py --config refdefault --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-baseline-method"
for v in dda1 ddalin; do
  export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so
  (timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_refdefault_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_refdefault_$v.json')); print('refdefault $v', round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['phases_ms_rank0'].items()}, d['parity']['status'] if d.get('parity') else None)"
done
