# round 2, call ab (4 GPUs): the default line (weak + strong leg) and C3 (strong) at N = 4
TAG=${1:-r2ab}
N=${2:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 --no-drt --supergrid 2>gpurun_out/${TAG}_n$N.err | grep '^{' | tail -1) > gpurun_out/${TAG}_bench_n$N.json
(timeout 200 $TR --master-port 29532 bench.py --gpus $N --config c3 --steps 5 --warmup 3 --no-e2e --no-baseline-method --no-drt 2>gpurun_out/${TAG}_c3.err | grep '^{' | tail -1) > gpurun_out/${TAG}_bench_c3_n$N.json
python - <<PY
import json
for f in ('n$N', 'c3_n$N'):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_%s.json' % f))
        print(f, 'step %.3f ms' % d['ms_per_step'], 'value %.1f' % d['value'], d.get('phases_ms_rank0'))
        print('   strong', d.get('strong')); print('   e2e', d.get('e2e') and d['e2e']['ms_per_step'], d.get('e2e_device_optimizer') and d['e2e_device_optimizer']['ms_per_step'])
    except Exception as e:
        print(f, 'failed', e)
PY
