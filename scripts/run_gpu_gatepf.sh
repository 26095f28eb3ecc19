mkdir -p gpurun_out
TAG=${TAG:-gatepf}
run() {
  env $2 timeout 300 python bench.py --no-e2e --no-cpu --no-also --model nacagat $3 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']), d['ms_per_step'], d['stages'], (d['parity'] or {}).get('ok'), d['clocks']['sm_mhz'])"
}
run pf1 MPO_GATE_L2PF=1 ""
run pf0 MPO_GATE_L2PF=0 "--no-parity"
run pf1b MPO_GATE_L2PF=1 "--no-parity"
run pf0b MPO_GATE_L2PF=0 "--no-parity"
