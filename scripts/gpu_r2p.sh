# round 2, call p: flattened pass 2 of the adjoint (RCV_FLAT_PASS2) — parity on the in-tree build, then A/B against the list walk
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c1.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
bash scripts/gpu_ab_variant.sh ${TAG} noflat
