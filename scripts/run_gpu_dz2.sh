# pipelined dz kernel: parity tests that cover the bag backward, then timing
mkdir -p gpurun_out
TAG=${TAG:-dz2}
echo "== quick check (old kernel: dW 1.203405e+00 db 1.128703e+00 dqk 8.143266e-01)"
timeout 90 python scripts/gpu_time_bwd.py 30 2>&1 | grep -v Warning || { echo "quick check failed / hung: stop"; exit 1; }
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest.log
for d in ${DBGS:-0 16 63}; do
  echo "== MPO_DZ_DEBUG=$d"
  MPO_DZ_DEBUG=$d timeout 120 python scripts/gpu_time_bwd.py 30 2>&1 | grep -v Warning
done | tee gpurun_out/${TAG}.log
timeout 300 python bench.py --no-e2e --no-cpu --no-also > gpurun_out/${TAG}_bench_mcat.json 2> gpurun_out/${TAG}_bench_mcat.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_mcat.json').read().strip().splitlines()[-1]);print('mcat', round(d['value']), d['ms_per_step'], d['stages'], d['parity'])"
timeout 300 python bench.py --model nacagat --no-e2e --no-cpu --no-also > gpurun_out/${TAG}_bench_nac.json 2> gpurun_out/${TAG}_bench_nac.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_nac.json').read().strip().splitlines()[-1]);print('nac', round(d['value']), d['ms_per_step'], d['stages'], d['parity'])"
