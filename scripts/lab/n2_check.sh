#!/bin/bash
TAG=${1:-n2}
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_pytest_multi.txt 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
