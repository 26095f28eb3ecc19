# round-end style check on one B200: parity tests, smoke, the default bench line (both arms), NaCAGaT line, launch lists
set -x
mkdir -p gpurun_out
TAG=${TAG:-fin2}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench_ref.json
timeout 400 python bench.py --model nacagat --no-e2e --no-cpu --no-also > gpurun_out/${TAG}_bench_nac.json 2> gpurun_out/${TAG}_bench_nac.err; echo "nac rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench_nac.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_mcat.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-also --no-parity > gpurun_out/${TAG}_ncu_mcat.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_nac.csv python bench.py --model nacagat --steps 2 --warmup 3 --no-e2e --no-cpu --no-also --no-parity > gpurun_out/${TAG}_ncu_nac.log 2>&1; echo "ncu rc=$?"
