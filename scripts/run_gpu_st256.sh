mkdir -p gpurun_out
TAG=${TAG:-st256}
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -x -q -k "nacagat or gate or ragged or large" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/${TAG}_pytest.log
run() {
  env $2 timeout 300 python bench.py --no-e2e --no-cpu --no-also --model nacagat $3 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']), d['ms_per_step'], d['stages'], (d['parity'] or {}).get('ok'), d['clocks']['sm_mhz'])"
}
run a A=1 ""
run b A=1 "--no-parity"
