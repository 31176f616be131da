#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"; grep -E "BAD|EXC|icp:|rror" gpurun_out/conv_cases.log | head
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 2>&1 | grep -E "dprob|rror|icp" | head -3
for i in 1 2; do
  ICP_NO_UNEVEN_N=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/EVEN /"
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/UNEVEN /"
done
ICP_NO_UNEVEN_N=1 timeout 300 python tools/gpu_bringup.py layer --name c17_1x1_768_704 --reps 10 --n 64 | grep LAYER | sed 's/^/EVEN /'
timeout 300 python tools/gpu_bringup.py layer --name c17_1x1_768_704 --reps 10 --n 64 | grep LAYER | sed "s/^/UNEVEN /"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
