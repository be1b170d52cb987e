# round 2, call y (8 GPUs): the named multi-GPU workloads of BASELINE.json — C3 (plume 256^3, 1024^2, 256 spp, strong), C4 (dense 512^3,
# 1024^2, 512 spp, strong, 2 GiB gradient all-reduce), C5 (the 200-iteration loop at 1024^2, 128 spp)
TAG=${1:-r2y}
N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 200 $TR --master-port 29521 bench.py --gpus $N --config c3 --steps 5 --warmup 3 --no-e2e --no-baseline-method --no-drt 2>gpurun_out/${TAG}_c3.err | grep '^{' | tail -1) > gpurun_out/${TAG}_bench_c3_n$N.json
(timeout 200 $TR --master-port 29522 bench.py --gpus $N --config c5 --steps 200 2>gpurun_out/${TAG}_c5.err | grep '^{' | tail -1) > gpurun_out/${TAG}_bench_c5_n$N.json
(timeout 400 $TR --master-port 29523 bench.py --gpus $N --config c4 --steps 3 --warmup 2 --no-e2e --no-baseline-method --no-drt 2>gpurun_out/${TAG}_c4.err | grep '^{' | tail -1) > gpurun_out/${TAG}_bench_c4_n$N.json
python - <<PY
import json
for f in ('c3', 'c5', 'c4'):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_%s_n$N.json' % f))
        print(f, 'step %.3f ms' % d['ms_per_step'], 'value %.1f' % d['value'], d.get('phases_ms_rank0'), d.get('stopwatch_ms'), d.get('loop_seconds'))
    except Exception as e:
        print(f, 'failed', e)
PY
tail -n 4 gpurun_out/${TAG}_c3.err gpurun_out/${TAG}_c5.err gpurun_out/${TAG}_c4.err | grep -v OMP_NUM | grep -v '^\*' | tail -12
