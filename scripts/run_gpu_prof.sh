# epilogue-only timing build of the forward bag kernel under ncu (source-level stall samples), then plain timings
set -x
mkdir -p gpurun_out
for d in 0 24 32; do MPO_FWD_DEBUG=$d timeout 120 python scripts/gpu_time_bag.py 32 fwd 2>&1 | tail -1; done | tee gpurun_out/s4b_decomp.log
MPO_FWD_DEBUG=24 timeout 300 ncu --set full --import-source on --clock-control none -k regex:bag_fwd_kernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_s4b_epi python scripts/gpu_time_bag.py 32 fwd > gpurun_out/s4b_ncu_epi.log 2>&1
MPO_FWD_DEBUG=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:bag_fwd_kernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_s4b_full python scripts/gpu_time_bag.py 32 fwd > gpurun_out/s4b_ncu_full.log 2>&1
ls -la gpurun_out
