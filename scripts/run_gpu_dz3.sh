# dz stage with three tile buffers and compact fp16 operands: quick check, parity tests, timing, bench lines
mkdir -p gpurun_out
TAG=${TAG:-dz3}
echo "== quick check (before: dW 1.203405e+00 db 1.128703e+00 dqk 8.143266e-01, bag_bwd 0.373 ms)"
timeout 90 python scripts/gpu_time_bwd.py 30 2>&1 | grep -v Warning || { echo "quick check failed / hung: stop"; exit 1; }
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest.log
run() {
  env $2 timeout 300 python bench.py --no-e2e --no-cpu --no-also $3 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']), d['ms_per_step'], d['stages'], d['parity'])"
}
run mcat A=1 ""
run nac A=1 "--model nacagat"
