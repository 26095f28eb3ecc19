# two-GPU check of the final build: multi-GPU tests, then both bench arms under torchrun as the driver launches them
set -x
mkdir -p gpurun_out
TAG=${TAG:-n2}
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_peer_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench_ref.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
