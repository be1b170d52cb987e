# round 2, call aj: ncu evidence of the kernels as they stand at the end of the round (sigma_t in the vertex record, warp-wide
# record scatter): launch list of one C2 step and a --set full capture of the adjoint (refreshes profiles/traffic.json)
TAG=${1:-r2aj}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-drt --supergrid --no-baseline-method"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vps_pool -s 1 -c 1 -o gpurun_out/${TAG}_vps $B > gpurun_out/${TAG}_ncu_vps.log 2>&1
ls -la gpurun_out/${TAG}*.ncu-rep
