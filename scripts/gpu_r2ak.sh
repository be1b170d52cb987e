# round 2, call ak: warp-wide record scatter with the next round's records requested ahead (coop2), and the same with the
# pass-2 vertex prefetch (pfv2), against r2af's version (coop)
TAG=${1:-r2ak}
mkdir -p gpurun_out
export RCV_B200_LIB=$PWD/build/variants/librcv_coop2.so
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c1.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --supergrid --no-baseline-method"
for v in coop coop2 pfv2 coop coop2 pfv2; do
  export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so
  (timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_$v.json')); print('$v', round(d['ms_per_step'],2), 'adjoint', round(d['phases_ms_rank0']['adjoint'],2), round(d['kernels']['volpathsimple_adjoint']['ms'],2))"
done
