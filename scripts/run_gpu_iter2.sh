# iteration check: parity tests touching the bag kernels, then bench lines (MCAT / NaCAGaT) and the NaCAGaT launch list
set -x
mkdir -p gpurun_out
TAG=${TAG:-it}
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --no-e2e --no-cpu --no-also > gpurun_out/${TAG}_bench_mcat.json 2> gpurun_out/${TAG}_bench_mcat.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/${TAG}_bench_mcat.json
timeout 300 python bench.py --model nacagat --no-e2e --no-cpu --no-also > gpurun_out/${TAG}_bench_nac.json 2> gpurun_out/${TAG}_bench_nac.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/${TAG}_bench_nac.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_nac.csv python bench.py --model nacagat --steps 2 --warmup 3 --no-e2e --no-cpu --no-also --no-parity > gpurun_out/${TAG}_ncu_nac.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_mcat.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-also --no-parity > gpurun_out/${TAG}_ncu_mcat.log 2>&1; echo "ncu rc=$?"
