set -x
mkdir -p gpurun_out
TAG=${TAG:-it}
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -x -q -k "nacagat" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/${TAG}_pytest.log
(timeout 100 python scripts/gpu_time_ge.py 4096; timeout 100 python scripts/gpu_time_ge.py 4096; MPO_GE_TC=0 timeout 100 python scripts/gpu_time_ge.py 4096; timeout 100 python scripts/gpu_time_ge.py 8192) > gpurun_out/${TAG}_ge_time.log 2>&1
cat gpurun_out/${TAG}_ge_time.log
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_ge4k_launches.csv python scripts/gpu_time_ge.py 4096 > gpurun_out/${TAG}_ge4k_ncu.log 2>&1; echo rc=$?
timeout 300 python bench.py --model nacagat --no-e2e --no-cpu --no-also > gpurun_out/${TAG}_bench_nac.json 2> gpurun_out/${TAG}_bench_nac.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench_nac.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_nac.csv python bench.py --model nacagat --steps 2 --warmup 3 --no-e2e --no-cpu --no-also --no-parity > gpurun_out/${TAG}_ncu_nac.log 2>&1; echo "ncu rc=$?"
