# eight-GPU bench of the final build, launched as the driver launches it
set -x
mkdir -p gpurun_out
TAG=${TAG:-n8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.err
