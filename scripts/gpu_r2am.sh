# round 2, call am: launch list of one reference-default iteration with the kernels at the end of the round
TAG=${1:-r2am}
mkdir -p gpurun_out
RD="python bench.py --config refdefault --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-baseline-method"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_refdefault_launches.csv $RD > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/${TAG}_refdefault_launches.csv')))
for i, r in enumerate(rows):
    if 'Kernel Name' in r: h = r; start = i; break
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
out = [(r[ki][:60], float(r[vi].replace(',', ''))) for r in rows[start + 1:] if len(r) > vi]
for k, v in out[-26:]: print('%9.1f us  %s' % (v / 1000, k))
PY
