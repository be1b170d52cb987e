# round 2, call al: DRT walk K tentative collisions at a time (RCV_DRT_AHEAD 1 / 2 / 4), reference-default line
TAG=${1:-r2al}
mkdir -p gpurun_out
export RCV_B200_LIB=$PWD/build/variants/librcv_drt4.so
timeout 300 python -m pytest tests -q -m gpu -x -k "drt or DRT" > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
B="python bench.py --config refdefault --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-baseline-method"
for v in drt1 drt2 drt4 drt1 drt4; do
  export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so
  (timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_refdefault_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_refdefault_$v.json')); print('refdefault $v', round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['phases_ms_rank0'].items()}, d['parity']['status'] if d.get('parity') else None)"
done
