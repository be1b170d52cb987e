# round 2, call g: full GPU suite on the restored tree, C1 (L2-resident) line + launch list + ncu of both C1 tracers, the reference-default line
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
(timeout 300 python bench.py --config c1 --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench_c1.err | tail -1) > gpurun_out/${TAG}_bench_c1.json
(timeout 600 python bench.py --config refdefault 2>gpurun_out/${TAG}_bench_refdefault.err | tail -1) > gpurun_out/${TAG}_bench_refdefault.json
C1="python bench.py --config c1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --no-baseline-method"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_c1_launches.csv $C1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_twostates_pool -s 3 -c 1 -o gpurun_out/${TAG}_c1_twostate $C1 > gpurun_out/${TAG}_ncu_ts.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_vps_pool -s 3 -c 1 -o gpurun_out/${TAG}_c1_vps $C1 > gpurun_out/${TAG}_ncu_vps.log 2>&1
for f in c1 refdefault; do python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_bench_$f.json'))
    print('$f', 'step %.3f ms' % d['ms_per_step'], 'value %.0f' % d['value'], 'roofline', {k: d['roofline'][k] for k in ('bound','kernel','achieved','peak','frac')})
    print('  phases', d['phases_ms_rank0']); print('  parity', d['parity'])
except Exception as e:
    print('$f failed', e)
PY
done
tail -3 gpurun_out/${TAG}_bench_c1.err gpurun_out/${TAG}_bench_refdefault.err
ls -la gpurun_out
