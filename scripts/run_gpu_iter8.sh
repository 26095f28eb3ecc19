set -x
mkdir -p gpurun_out
TAG=${TAG:-it}
for v in 0 1; do
  MPO_DW_DZ_NORMAL=$v timeout 300 python bench.py --no-e2e --no-cpu --no-also > gpurun_out/${TAG}_bench_mcat_dzn$v.json 2> gpurun_out/${TAG}_bench_mcat_dzn$v.err; echo "bench rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_mcat_dzn$v.json').read().strip().splitlines()[-1]);print('dz_normal=$v', round(d['value']), d['ms_per_step'], d['stages'], d['clocks'])"
done
for v in 0 1; do
  MPO_DW_DZ_NORMAL=$v timeout 300 python bench.py --steps 2000 --no-e2e --no-cpu --no-also --no-parity > gpurun_out/${TAG}_bench_mcat_long_dzn$v.json 2>/dev/null
  python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_mcat_long_dzn$v.json').read().strip().splitlines()[-1]);print('long dz_normal=$v', round(d['value']), d['ms_per_step'], d['stages'], d['clocks'])"
done
MPO_DW_DZ_NORMAL=1 timeout 300 python bench.py --model nacagat --no-e2e --no-cpu --no-also > gpurun_out/${TAG}_bench_nac.json 2> gpurun_out/${TAG}_bench_nac.err
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_nac.json').read().strip().splitlines()[-1]);print('nac', round(d['value']), d['ms_per_step'], d['stages'], d['clocks'])"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused_cluster or wide_snn or split_adam" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/${TAG}_pytest.log
