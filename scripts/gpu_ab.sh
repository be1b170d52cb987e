# usage: bash scripts/gpu_ab.sh <tag> "<bench args>" <variant|default> ...
TAG=$1; ARGS=$2; shift 2
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e $ARGS"
for v in "$@"; do
  if [ "$v" = "default" ]; then (timeout 300 $B 2>&1 | tail -1) > gpurun_out/${TAG}_default.json
  else (RCV_B200_LIB=$PWD/build/variants/librcv_$v.so timeout 300 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$v.json; fi
done
