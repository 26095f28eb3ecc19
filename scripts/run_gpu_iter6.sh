set -x
mkdir -p gpurun_out
TAG=${TAG:-it}
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ge_nacagat" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/${TAG}_pytest.log
(timeout 100 python scripts/gpu_time_ge.py 4096; MPO_GE_TC=0 timeout 100 python scripts/gpu_time_ge.py 4096; timeout 100 python scripts/gpu_time_ge.py 8192; timeout 200 python scripts/gpu_time_ge.py 16384) > gpurun_out/${TAG}_ge_time.log 2>&1
grep GE- gpurun_out/${TAG}_ge_time.log
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_ge16k_launches.csv python scripts/gpu_time_ge.py 16384 > gpurun_out/${TAG}_ge16k_ncu.log 2>&1; echo rc=$?
