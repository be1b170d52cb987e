# usage: bash scripts/gpu_drt.sh <tag> [variant ...]   (run under gpurun): GPU tests, smoke, then bench.py --use-drt with
# the in-tree library and with each build/variants/librcv_<variant>.so; last, the ncu launch list of one DRT step
TAG=$1; shift
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --supergrid --use-drt"
for v in default "$@"; do
  if [ "$v" = "default" ]; then unset RCV_B200_LIB; else export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so; fi
  (timeout 200 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$v.json
done
unset RCV_B200_LIB
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --supergrid --use-drt > /dev/null 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_smoke.log
for v in default "$@"; do python -c "
import json
try:
    d=json.load(open('gpurun_out/${TAG}_$v.json')); print('$v', d['ms_per_step'], d['value'], d.get('phases_ms_rank0'), d['e2e'], d['gpu_launches'])
except Exception as e: print('$v', 'failed', e)
"; done
