mkdir -p gpurun_out
TAG=${TAG:-fwdvar}
run() {
  env $2 timeout 300 python bench.py --no-e2e --no-cpu --no-also --no-parity $3 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']), d['ms_per_step'], d['stages'], d['clocks']['sm_mhz'])"
}
run base A=1 ""
run pair MPO_FWD_PAIR=1 ""
run base2 A=1 ""
run pair2 MPO_FWD_PAIR=1 ""
run cl1 MPO_FWD_CLUSTER=1 ""
run cl4 MPO_FWD_CLUSTER=4 ""
