TAG=${1:-rX}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_twostates_pool -s 1 -c 1 -o gpurun_out/${TAG}_twostate \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --tracer 3 > gpurun_out/${TAG}_ncu_ts_stdout.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vps_pool -s 1 -c 1 -o gpurun_out/${TAG}_vps \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --tracer 3 > gpurun_out/${TAG}_ncu_vps_stdout.log 2>&1
ls -la gpurun_out
