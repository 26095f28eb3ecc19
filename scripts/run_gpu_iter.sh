set -x
mkdir -p gpurun_out
for d in 0 256 512 768 0 256; do MPO_FWD_DEBUG=$d timeout 120 python scripts/gpu_time_bag.py 32 fwd 2>&1 | tail -1; done | tee gpurun_out/iter_decomp.log
