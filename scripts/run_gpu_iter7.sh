set -x
mkdir -p gpurun_out
TAG=${TAG:-it}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
(timeout 100 python scripts/gpu_time_ge.py 4096; MPO_GE_TC=0 timeout 100 python scripts/gpu_time_ge.py 4096; timeout 100 python scripts/gpu_time_ge.py 8192; timeout 200 python scripts/gpu_time_ge.py 16384) > gpurun_out/${TAG}_ge_time.log 2>&1
grep GE- gpurun_out/${TAG}_ge_time.log
timeout 300 python bench.py --no-e2e --no-cpu --no-also > gpurun_out/${TAG}_bench_mcat.json 2> gpurun_out/${TAG}_bench_mcat.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench_mcat.json
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_ge16k_launches.csv python scripts/gpu_time_ge.py 16384 > gpurun_out/${TAG}_ge16k_ncu.log 2>&1; echo rc=$?
