# round 2, call r: ncu of the surface machine after the restructuring (c1box, both kernels)
TAG=${1:-r2r}
mkdir -p gpurun_out
C1B="python bench.py --config c1box --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-baseline-method"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_surface_wavefront -s 4 -c 2 -o gpurun_out/${TAG}_c1box $C1B > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}_c1box.ncu-rep
