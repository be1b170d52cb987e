# usage: bash scripts/gpu_ab_variant.sh <tag> <variant>   (run under gpurun): the parity tests on build/variants/librcv_<variant>.so,
# then a short bench of the variant and of the in-tree library
TAG=$1; V=$2
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --supergrid"
export RCV_B200_LIB=$PWD/build/variants/librcv_$V.so
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/${TAG}_${V}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_${V}_pytest.log
(timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$V.json
unset RCV_B200_LIB
(timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_default.json
for v in $V default; do python -c "
import json
d=json.load(open('gpurun_out/${TAG}_$v.json')); print('$v', d['ms_per_step'], d['phases_ms_rank0']['adjoint'], d['kernels']['volpathsimple_adjoint']['ms'], d['gpu_launches'])"; done
