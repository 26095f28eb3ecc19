# round-2 (second session) evidence: stand-alone block tests, ncu --set full of the new / changed kernels, launch lists
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_blocks_gpu.py -m gpu -x -q > gpurun_out/r2c_blocks_pytest.log 2>&1; echo "blocks pytest rc=$?"
tail -3 gpurun_out/r2c_blocks_pytest.log
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-also --no-parity"
# MCAT step: 8 matching launches per step (snn2_fwd x2, bag_fwd, path x2, dz, dw, snn2_bwd) -> skip 3 steps, take 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"snn2_fwd_kernel|snn2_bwd_kernel|bag_fwd_kernel|bag_bwd_dz_kernel|bag_bwd_dw_kernel|path_kernel" -s 24 -c 8 -f -o gpurun_out/r2c_full_mcat $BENCH > gpurun_out/r2c_full_mcat.log 2>&1; echo "ncu mcat rc=$?"
# NaCAGaT step: gate, dz<1>, dz<2>, dhk
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"bag_gate_kernel|bag_dhk_kernel|bag_bwd_dz_kernel" -s 12 -c 4 -f -o gpurun_out/r2c_full_nac $BENCH --model nacagat > gpurun_out/r2c_full_nac.log 2>&1; echo "ncu nac rc=$?"
# GE-NaCAGaT at 4 096 patches: the tensor-core GEMM in its three shapes, the fused soft-max backward, the dropout split
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gemm_tc_kernel|row_softmax_bwd_pair_kernel|split_bf16_kernel" -s 600 -c 12 -f -o gpurun_out/r2c_full_ge python scripts/gpu_time_ge.py 4096 > gpurun_out/r2c_full_ge.log 2>&1; echo "ncu ge rc=$?"
ls -la gpurun_out/r2c_*
