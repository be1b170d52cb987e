# round 2, call n (2 GPUs): the default line at N = 2 (weak) with its strong-scaling leg, c3 strong, the c5 loop under torchrun
TAG=${1:-r2n}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 400 $TR --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/${TAG}_bench_n2.err | grep '^{' | tail -1) > gpurun_out/${TAG}_bench_n2.json
(timeout 300 $TR --master-port 29512 bench.py --gpus 2 --config c3 --steps 3 --warmup 3 --no-e2e --no-baseline-method --no-drt 2>gpurun_out/${TAG}_bench_c3_n2.err | grep '^{' | tail -1) > gpurun_out/${TAG}_bench_c3_n2.json
(timeout 300 $TR --master-port 29513 bench.py --gpus 2 --config c5 --steps 40 2>gpurun_out/${TAG}_bench_c5_n2.err | grep '^{' | tail -1) > gpurun_out/${TAG}_bench_c5_n2.json
python - <<PY
import json
for f in ('n2', 'c3_n2', 'c5_n2'):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_%s.json' % f))
        print(f, 'step %.3f ms' % d['ms_per_step'], 'value %.1f' % d['value'], d.get('phases_ms_rank0'), 'strong', d.get('strong'), d.get('stopwatch_ms'))
        print('  e2e', d.get('e2e'))
    except Exception as e:
        print(f, 'failed', e)
PY
tail -n 5 gpurun_out/${TAG}_bench_n2.err gpurun_out/${TAG}_bench_c3_n2.err gpurun_out/${TAG}_bench_c5_n2.err
