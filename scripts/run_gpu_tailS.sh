# fused tail with one vs two slides per cluster (two: 6-slot weight ring) at 32 and 128 slides per step
mkdir -p gpurun_out
TAG=${TAG:-tailS}
run() {  # name, env, extra args
  env $2 timeout 300 python bench.py --no-e2e --no-cpu --no-also $3 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']), d['ms_per_step'], d['stages'], (d['parity'] or {}).get('ok'), (d['parity'] or {}).get('grad_worst_rel_err'))"
}
run mcat_S1 MPO_TAIL_FUSED_S=1 ""
run mcat_S2 MPO_TAIL_FUSED_S=2 ""
run nac_S1 MPO_TAIL_FUSED_S=1 "--model nacagat"
run nac_S2 MPO_TAIL_FUSED_S=2 "--model nacagat"
run mcat_B128_S2 MPO_TAIL_FUSED_S=2 "--batch 128 --no-parity"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused or batch or graph" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
