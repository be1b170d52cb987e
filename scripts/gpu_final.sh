# what the driver does at round end, on the in-tree build: GPU test suite, smoke(), the default bench line and the reference arm
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/${TAG}_ref.err | tail -1) > gpurun_out/${TAG}_bench_reference_arm.json
(timeout 600 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_n1.json
python -c "
import json
r=json.load(open('gpurun_out/${TAG}_bench_reference_arm.json')); d=json.load(open('gpurun_out/${TAG}_bench_n1.json'))
print('reference arm', r['value'], r['cpu_baseline'], r['config']==d['config'])
print('ours', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e_device_optimizer']['ms_per_step'], 'launches', d['gpu_launches'], 'parity', d['parity']['pass'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])"
tail -n 2 gpurun_out/${TAG}_ref.err gpurun_out/${TAG}_bench.err
