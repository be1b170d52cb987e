# usage: bash scripts/gpu_profile.sh <tag>   (run under gpurun; writes gpurun_out/<tag>_*)
TAG=${1:-rX}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${TAG}_pytest.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
(timeout 300 $B 2>&1 | tail -1) > gpurun_out/${TAG}_bench.json
# launch list (cold-cache, serialised: compare shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_launch_stdout.log 2>&1
# full capture of the two tracing kernels (default schedule: pool)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_twostates_pool -s 1 -c 1 -o gpurun_out/${TAG}_twostate \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_ts_stdout.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vps_pool -s 1 -c 1 -o gpurun_out/${TAG}_vps \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_vps_stdout.log 2>&1
tail -2 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json | cut -c1-300; ls -la gpurun_out
