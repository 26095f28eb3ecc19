mkdir -p gpurun_out
TAG=${TAG:-tail4}
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused_cluster" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
run() {
  env $2 timeout 300 python bench.py --no-e2e --no-cpu --no-also $3 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']), d['ms_per_step'], d['stages'], (d['parity'] or {}).get('ok'), d['clocks']['sm_mhz'])"
}
run mcat_S1 MPO_TAIL_FUSED_S=1 "--no-parity"
run mcat_S2 MPO_TAIL_FUSED_S=2 "--no-parity"
run mcat_S1b MPO_TAIL_FUSED_S=1 "--no-parity"
run mcat_S2b MPO_TAIL_FUSED_S=2 "--no-parity"
run nac_S2 MPO_TAIL_FUSED_S=2 "--model nacagat --no-parity"
