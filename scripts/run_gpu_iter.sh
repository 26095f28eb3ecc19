set -x
mkdir -p gpurun_out
for pr in 1 0 1; do MPO_FWD_PAIR=$pr timeout 120 python scripts/gpu_clock_probe.py 2>&1 | tail -1; done | tee gpurun_out/iter_clock2.log
MPO_FWD_PAIR=1 MPO_FWD_DEBUG=32 timeout 120 python scripts/gpu_clock_probe.py 2>&1 | tail -1 | tee -a gpurun_out/iter_clock2.log
