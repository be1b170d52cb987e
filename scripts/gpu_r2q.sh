# round 2, call q: is the adjoint DRAM-bound on the gradient read-modify-writes now?  d sigma_t reductions on the 64 MiB voxel grid
# (L2-sized; RCV_GRAD_BRICKS=0), with and without the L2 persistence window, against the gradient bricks
TAG=${1:-r2q}
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-drt --supergrid --no-baseline-method"
run() {  # name, lib, env
  ( export $3; [ -n "$2" ] && export RCV_B200_LIB=$PWD/build/variants/librcv_$2.so; (timeout 100 $B 2>&1 | tail -1) > gpurun_out/${TAG}_$1.json )
  python -c "
import json
d=json.load(open('gpurun_out/${TAG}_$1.json')); print('$1', round(d['ms_per_step'],2), 'adjoint', round(d['phases_ms_rank0']['adjoint'],2), round(d['kernels']['volpathsimple_adjoint']['ms'],2))"
}
export RCV_B200_LIB=$PWD/build/variants/librcv_voxred.so
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/${TAG}_voxred_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_voxred_pytest.log
unset RCV_B200_LIB
run default "" X=1
run voxred voxred X=1
run voxred_persist voxred RCV_L2_PERSIST=1
run default_persist "" RCV_L2_PERSIST=1
