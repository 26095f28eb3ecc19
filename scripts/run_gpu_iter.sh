set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/iter_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/iter_pytest.log
timeout 300 python scripts/gpu_profile_module_call.py 16384 > gpurun_out/modprof2.log 2>&1
grep "ms per call" gpurun_out/modprof2.log
