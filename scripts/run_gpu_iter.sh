set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/iter_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/iter_pytest.log
MPO_FWD_DEBUG=128 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/iter_pytest128.log 2>&1; echo "pytest128 rc=$?"
tail -2 gpurun_out/iter_pytest128.log
for d in 0 128 0 128 144 16; do MPO_FWD_DEBUG=$d timeout 120 python scripts/gpu_time_bag.py 32 fwd 2>&1 | tail -1; done | tee gpurun_out/iter_decomp.log
