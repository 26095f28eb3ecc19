# round-end check on the GPU box: parity tests, smoke, bench line, ncu launch list of the same bench command
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/fin_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/fin_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/fin_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/fin_smoke.log
timeout 900 python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench rc=$?"
cat gpurun_out/fin_bench.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-also > gpurun_out/fin_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python scripts/gpu_profile_module_call.py 16384 > gpurun_out/fin_modprof.log 2>&1; grep "ms per call" gpurun_out/fin_modprof.log
