# C1 GPU tests, the L2 gather microbenchmark, the full default bench line and the C1 line
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_c1.py -q -m gpu -x -s > gpurun_out/${TAG}_c1_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_c1_pytest.log
timeout 300 ./build/gather_bw l2 > gpurun_out/${TAG}_gather_l2.jsonl 2> gpurun_out/${TAG}_gather_l2.err; tail -1 gpurun_out/${TAG}_gather_l2.jsonl
(timeout 600 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_c2.json
(timeout 300 python bench.py --config c1 --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench_c1.err | tail -1) > gpurun_out/${TAG}_bench_c1.json
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_ref.json
for f in c2 c1; do python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_bench_$f.json'))
    print('$f', 'step %.3f ms' % d['ms_per_step'], 'value %.0f' % d['value'], 'roofline', {k: d['roofline'][k] for k in ('bound','kernel','achieved','peak','frac')})
    print('  e2e', d['e2e'] and d['e2e']['ms_per_step'], 'e2e_dev', d['e2e_device_optimizer'] and d['e2e_device_optimizer']['ms_per_step'], 'baseline', d['baseline_method'])
    print('  parity', d['parity'])
    print('  cpu', d['cpu_baseline'])
except Exception as e:
    print('$f failed', e)
PY
done
tail -3 gpurun_out/${TAG}_bench.err gpurun_out/${TAG}_bench_c1.err
