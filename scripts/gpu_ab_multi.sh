# usage: bash scripts/gpu_ab_multi.sh <tag> "<bench args>" <variant|default> ...   (run under gpurun)
# per variant (build/variants/librcv_<v>.so, or the in-tree library for `default`): the GPU parity tests, then a short bench
TAG=$1; ARGS=$2; shift 2
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --supergrid $ARGS"
for v in "$@"; do
  if [ "$v" = "default" ]; then unset RCV_B200_LIB; else export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/${TAG}_${v}_pytest.log 2>&1; echo "$v: $(tail -1 gpurun_out/${TAG}_${v}_pytest.log)"
  (timeout 200 $B 2>gpurun_out/${TAG}_${v}.err | tail -1) > gpurun_out/${TAG}_$v.json
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_$v.json'))
    k = d['kernels']
    print('$v', 'step %.2f ms' % d['ms_per_step'], 'two-state %.2f' % k['twostate_primal']['ms'], 'adjoint %.2f' % k['volpathsimple_adjoint']['ms'],
          {a: round(b, 2) for a, b in d['phases_ms_rank0'].items()})
except Exception as e:
    print('$v', 'failed', e)
PY
done
