mkdir -p gpurun_out
TAG=${TAG:-tail5}
echo "== quick check"; timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused_cluster" > gpurun_out/${TAG}_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -3 gpurun_out/${TAG}_pytest.log
[ $rc = 0 ] || exit 1
run() {
  env $2 timeout 300 python bench.py --no-e2e --no-cpu --no-also $3 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']), d['ms_per_step'], d['stages'], (d['parity'] or {}).get('ok'), d['clocks']['sm_mhz'])"
}
run mcat_tma1 "MPO_TAIL_TMA=1 MPO_TAIL_VERBOSE=1" ""
grep "\[mpo\]" gpurun_out/${TAG}_mcat_tma1.err; run mcat_tma0 MPO_TAIL_TMA=0 "--no-parity"
run mcat_tma1b MPO_TAIL_TMA=1 "--no-parity"
run mcat_tma0b MPO_TAIL_TMA=0 "--no-parity"
run nac_tma1 MPO_TAIL_TMA=1 "--model nacagat"
run nac_tma0 MPO_TAIL_TMA=0 "--model nacagat --no-parity"
bash scripts/run_gpu_tailprof.sh > gpurun_out/${TAG}_prof_build.log 2>&1
grep -A8 "mcat.*pass B=32" gpurun_out/tailprof.log | head -24
