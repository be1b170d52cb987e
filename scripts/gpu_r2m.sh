# round 2, call m: the c2box line (BASELINE configs[1] as worded), and the full loop (c5) on one GPU for 40 iterations
TAG=${1:-r2m}
mkdir -p gpurun_out
(timeout 500 python bench.py --config c2box --steps 3 --warmup 3 --cpu-sample-spp 1 2>gpurun_out/${TAG}_bench_c2box.err | tail -1) > gpurun_out/${TAG}_bench_c2box.json
(timeout 300 python bench.py --config c5 --steps 40 2>gpurun_out/${TAG}_bench_c5_n1.err | tail -1) > gpurun_out/${TAG}_bench_c5_n1.json
python - <<PY
import json
for f in ('c2box', 'c5_n1'):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_%s.json' % f))
        print(f, 'step %.3f ms' % d['ms_per_step'], 'value %.1f' % d['value'], d.get('phases_ms_rank0'), d.get('stopwatch_ms'), d.get('loss_first_last'))
        print('  parity', d.get('parity')); print('  cpu', d.get('cpu_baseline')); print('  e2e', d.get('e2e'))
    except Exception as e:
        print(f, 'failed', e)
PY
tail -n 4 gpurun_out/${TAG}_bench_c2box.err gpurun_out/${TAG}_bench_c5_n1.err
