# quick check of the in-tree build: parity tests, smoke, a short bench; then the forward kernel under ncu, single vs pair tiles
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/iter_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/iter_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --no-also --no-cpu > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/iter_bench.json
for pr in 0 1; do MPO_FWD_PAIR=$pr timeout 300 ncu --set full --clock-control none -k regex:bag_fwd_kernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_s4_pair$pr python scripts/gpu_time_bag.py 32 fwd > gpurun_out/iter_ncu_pair$pr.log 2>&1; done
ls -la gpurun_out/*.ncu-rep
