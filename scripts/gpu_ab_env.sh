# A/B of the L2 options: default / persist window / streaming bricks / both
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e"
(timeout 300 $B 2>&1 | tail -1) > gpurun_out/$1_base.json
(RCV_L2_PERSIST=1 timeout 300 $B 2>&1 | tail -1) > gpurun_out/$1_persist.json
(RCV_B200_LIB=$PWD/build/variants/librcv_cs.so timeout 300 $B 2>&1 | tail -1) > gpurun_out/$1_cs.json
(RCV_L2_PERSIST=1 RCV_B200_LIB=$PWD/build/variants/librcv_cs.so timeout 300 $B 2>&1 | tail -1) > gpurun_out/$1_cs_persist.json
