# round 2, call ae: unroll of the four transmittance-term sites of a pass-2 vertex (RCV_TRANS_UNROLL 1 / 2 / 4)
TAG=${1:-r2ae}
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --supergrid --no-baseline-method"
for v in default tu2 tu4 default tu2 tu4; do
  if [ "$v" = "default" ]; then unset RCV_B200_LIB; else export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so; fi
  (timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_$v.json')); print('$v', round(d['ms_per_step'],2), 'adjoint', round(d['phases_ms_rank0']['adjoint'],2), round(d['kernels']['volpathsimple_adjoint']['ms'],2), d['parity']['status'] if d.get('parity') else None)"
done
