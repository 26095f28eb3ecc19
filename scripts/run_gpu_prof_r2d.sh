# final build of round 2: ncu --set full of the kernels that changed in the third session (dz stage, path kernels) inside the bench step
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-also --no-parity"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"bag_fwd_kernel|bag_bwd_dz_kernel|bag_bwd_dw_kernel|path_kernel" -s 15 -c 5 -f -o gpurun_out/r2d_full_mcat $BENCH > gpurun_out/r2d_full_mcat.log 2>&1; echo "ncu mcat rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"bag_bwd_dz_kernel" -s 6 -c 2 -f -o gpurun_out/r2d_full_nac $BENCH --model nacagat > gpurun_out/r2d_full_nac.log 2>&1; echo "ncu nac rc=$?"
ls -la gpurun_out/r2d_full_*
