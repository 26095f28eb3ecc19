# timing decomposition of bag_bwd_dz_kernel<0> with its MPO_DZ_DEBUG switches (results are wrong when a switch is set)
mkdir -p gpurun_out
TAG=${TAG:-dzdbg}
for d in ${DBGS:-0 1 2 3 4 8 16 32 11 15 27 63}; do
  echo "== MPO_DZ_DEBUG=$d"
  MPO_DZ_DEBUG=$d timeout 120 python scripts/gpu_time_bwd.py 30 2>&1 | grep -v Warning
done | tee gpurun_out/${TAG}.log
