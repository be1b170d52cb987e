# usage: bash scripts/gpu_variants.sh <tag> <variant> [<variant>...]   (run under gpurun)
TAG=$1; shift
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
for v in "$@"; do
  (RCV_B200_LIB=$PWD/build/variants/librcv_$v.so timeout 300 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$v.json
done
