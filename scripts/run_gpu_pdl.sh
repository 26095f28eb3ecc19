# programmatic dependent launch for the kernels of the fused step: parity tests, then bench with and without
mkdir -p gpurun_out
TAG=${TAG:-pdl}
echo "== quick check"; timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused_cluster or graph or batch" > gpurun_out/${TAG}_pytest_q.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -3 gpurun_out/${TAG}_pytest_q.log
[ $rc = 0 ] || exit 1
run() {
  env $2 timeout 300 python bench.py --no-e2e --no-cpu --no-also $3 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']), d['ms_per_step'], d['stages'], (d['parity'] or {}).get('ok'), d['clocks']['sm_mhz'])"
}
run mcat_pdl1 MPO_STEP_PDL=1 ""
run mcat_pdl0 MPO_STEP_PDL=0 "--no-parity"
run mcat_pdl1b MPO_STEP_PDL=1 "--no-parity"
run mcat_pdl0b MPO_STEP_PDL=0 "--no-parity"
run nac_pdl1 MPO_STEP_PDL=1 "--model nacagat"
run nac_pdl0 MPO_STEP_PDL=0 "--model nacagat --no-parity"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "full pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
