# round 2, call w: ncu evidence of the FINAL kernels at C2: launch list of one step, --set full captures of the adjoint and the two-state
# tracer (refreshes profiles/traffic.json through scripts/summarize_profiles.py), plus the default bench line
TAG=${1:-r2w}
mkdir -p gpurun_out
(timeout 400 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench.json
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-drt --supergrid --no-baseline-method"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vps_pool -s 1 -c 1 -o gpurun_out/${TAG}_vps $B > gpurun_out/${TAG}_ncu_vps.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_twostates_pool -s 1 -c 1 -o gpurun_out/${TAG}_twostate $B > gpurun_out/${TAG}_ncu_ts.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench.json')); print('c2', d['ms_per_step'], d['value'], d['phases_ms_rank0'], d['roofline'], 'e2e', d['e2e']['ms_per_step'], d['e2e_device_optimizer']['ms_per_step'], d['parity']['pass'])"
ls -la gpurun_out/${TAG}*.ncu-rep
