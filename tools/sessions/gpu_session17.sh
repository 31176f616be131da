#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench2.log 2>&1
echo "rc=$?"; tail -n 5 gpurun_out/bench2.log | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench2_ref.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/bench2_ref.log | cut -c1-400
