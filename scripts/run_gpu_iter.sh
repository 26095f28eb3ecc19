set -x
mkdir -p gpurun_out
for cfg in "0 0" "0 256" "1 0" "1 256" "0 256" "0 0"; do set -- $cfg; MPO_FWD_PAIR=$1 MPO_FWD_DEBUG=$2 timeout 100 python scripts/gpu_time_bag.py 32 fwd 2>&1 | tail -1 | sed "s/^/pair=$1 /"; done | tee gpurun_out/iter_early.log
