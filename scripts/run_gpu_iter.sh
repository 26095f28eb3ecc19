# quick check of the in-tree build: parity tests, smoke, a short bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/iter_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/iter_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/iter_bench.json
