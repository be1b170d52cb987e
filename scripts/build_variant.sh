# usage: bash scripts/build_variant.sh <name> "<extra nvcc flags>"   -> build/variants/librcv_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
C=recursive_control_variates_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared $2 \
  -o build/variants/librcv_$1.so $C/rcv_api.cu $C/rcv_trace.cu $C/rcv_surface.cu $C/rcv_pixel.cu 2>&1 | grep -i "error" || true
ls -la build/variants/librcv_$1.so
