# round 2, call s: occupancy of the surface machine (3 / 4 / 5 CTAs of 128 threads per SM), c1box
TAG=${1:-r2s}
mkdir -p gpurun_out
B="python bench.py --config c1box --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --no-baseline-method"
for v in default sf4 sf5; do
  if [ "$v" = "default" ]; then unset RCV_B200_LIB; else export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so; fi
  (timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_$v.json')); print('$v', round(d['ms_per_step'],2), d['phases_ms_rank0'])"
done
export RCV_B200_LIB=$PWD/build/variants/librcv_sf4.so
timeout 300 python -m pytest tests/test_surfaces.py -q -m gpu -x 2>&1 | tail -1
unset RCV_B200_LIB
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_surfaces.py -q -m gpu -k golden 2>&1 | tail -1
RD="python bench.py --config refdefault --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-baseline-method"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_refdefault_launches.csv $RD > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/${TAG}_refdefault_launches.csv')))
for i, r in enumerate(rows):
    if 'Kernel Name' in r: h = r; start = i; break
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
out = [(r[ki][:60], float(r[vi].replace(',', ''))) for r in rows[start + 1:] if len(r) > vi]
for k, v in out[-22:]: print('%9.1f us  %s' % (v / 1000, k))
PY
