# usage: bash scripts/gpu_round2_first.sh <tag>   (run under gpurun, one GPU, ~6 min of box time)
# Pays the measurement debts DESIGN.md §8 lists at the end of round 1:
#   1. GPU tests + smoke() on the in-tree library
#   2. default bench (the JSON line for profiles/<tag>_bench_n1.json)
#   3. A/B of the variants built beforehand with scripts/build_variant.sh, e.g.  va1 "-DRCV_VERTEX_ALBEDO=1"
#   4. launch list of one step and ONE `ncu --set full` capture of the adjoint (refreshes roofline.traffic: read it with
#      `ncu -i gpurun_out/<tag>_vps.ncu-rep --page raw --csv`, then update profiles/traffic.json via scripts/summarize_profiles.py)
TAG=${1:-r2a}; shift
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(timeout 400 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1) > gpurun_out/${TAG}_bench_n1.json
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --supergrid"
for v in "$@"; do
  RCV_B200_LIB=$PWD/build/variants/librcv_$v.so timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/${TAG}_${v}_pytest.log 2>&1
  tail -1 gpurun_out/${TAG}_${v}_pytest.log
  (RCV_B200_LIB=$PWD/build/variants/librcv_$v.so timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$v.json
done
(timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_default_short.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-drt --supergrid > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vps_pool -s 1 -c 1 -o gpurun_out/${TAG}_vps \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-drt --supergrid > gpurun_out/${TAG}_ncu_vps_stdout.log 2>&1
for f in gpurun_out/${TAG}_bench_n1.json gpurun_out/${TAG}_default_short.json $(for v in "$@"; do echo gpurun_out/${TAG}_$v.json; done); do python -c "
import json
try:
    d=json.load(open('$f')); print('$f', d['ms_per_step'], d['value'], d['phases_ms_rank0'])
except Exception as e: print('$f', 'failed', e)
"; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_twostates_pool -s 1 -c 1 -o gpurun_out/${TAG}_ts \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-drt --supergrid > gpurun_out/${TAG}_ncu_ts_stdout.log 2>&1
ls -la gpurun_out/*.ncu-rep
