# fused tail change: parity (fused vs per-op tail, graph, batch), both bench lines, then the phase timing of a prof build
mkdir -p gpurun_out
TAG=${TAG:-tail2}
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused or batch or graph or wide_snn or nacagat or mcat" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
run() {
  env $2 timeout 300 python bench.py --no-e2e --no-cpu --no-also $3 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_$1.json').read().strip().splitlines()[-1]);print('$1', round(d['value']), d['ms_per_step'], d['stages'], (d['parity'] or {}).get('ok'), (d['parity'] or {}).get('grad_worst_rel_err'))"
}
run mcat A=1 ""
run nac A=1 "--model nacagat"
run mcat_B128 A=1 "--batch 128 --no-parity"
bash scripts/run_gpu_tailprof.sh > gpurun_out/${TAG}_prof_build.log 2>&1
grep -A8 "pass B=32" gpurun_out/tailprof.log | head -40
