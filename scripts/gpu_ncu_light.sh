# usage: bash scripts/gpu_ncu_light.sh <tag> <kernel-regex> <variant|default> ...   (run under gpurun)
# per variant: a short bench, then one light ncu capture (scheduler / warp-state / memory sections) of the named kernel
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
for v in "$@"; do
  if [ "$v" = "default" ]; then unset RCV_B200_LIB; else export RCV_B200_LIB=$PWD/build/variants/librcv_$v.so; fi
  (timeout 300 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$v.json
  timeout 600 ncu --clock-control none -k regex:$KRE -s 2 -c 1 --section SchedulerStats --section WarpStateStats \
     --section InstructionStats --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
     --metrics smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex.sum,l1tex__t_sector_hit_rate.pct \
     --csv --page raw --log-file gpurun_out/${TAG}_${v}_ncu.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
done
