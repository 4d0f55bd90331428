#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): headline bench and BASELINE config 5 (inner product) on N GPUs.  Usage: tools/gpu_n8.sh <tag> <N>
TAG=${1:-n8}; N=${2:-8}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json; tail -2 $O/${TAG}_bench.err
timeout 900 $TR --master-port 29512 tools/inner_product.py --key-bits 1024 --length 65536 --slots 8 > $O/${TAG}_inner.json 2> $O/${TAG}_inner.err
cat $O/${TAG}_inner.json; tail -2 $O/${TAG}_inner.err
