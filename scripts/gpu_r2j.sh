# round 2, call j: where the surface machine spends its time (ncu, c1box), C1 with majorant supergrids (the tail), a bounded c2box line
TAG=${1:-r2j}
mkdir -p gpurun_out
C1B="python bench.py --config c1box --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-baseline-method"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_surface_wavefront -s 2 -c 1 -o gpurun_out/${TAG}_c1box_ts $C1B > gpurun_out/${TAG}_ncu_ts.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_c1box_launches.csv $C1B > /dev/null 2>&1
(timeout 300 python bench.py --config c1 --steps 20 --warmup 5 --supergrid 4 8 16 2>gpurun_out/${TAG}_bench_c1.err | tail -1) > gpurun_out/${TAG}_bench_c1.json
(timeout 800 python bench.py --config c2box --steps 2 --warmup 1 --no-e2e --no-baseline-method --cpu-sample-spp 1 2>gpurun_out/${TAG}_bench_c2box.err | tail -1) > gpurun_out/${TAG}_bench_c2box.json
python - <<PY
import json
for f in ('c1', 'c2box'):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_%s.json' % f))
        print(f, 'step %.3f ms' % d['ms_per_step'], 'value %.0f' % d['value'], d['phases_ms_rank0'], [(s['factor'], round(s['ms_per_step'], 3), s['twostate_collisions']) for s in d['supergrid']])
        print('  parity', d['parity']); print('  cpu', d['cpu_baseline'])
    except Exception as e:
        print(f, 'failed', e)
PY
tail -n 3 gpurun_out/${TAG}_bench_c1.err gpurun_out/${TAG}_bench_c2box.err
