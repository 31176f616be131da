#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"; grep -E "BAD|EXC|icp:" gpurun_out/conv_cases.log | head
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 | grep dprob
for i in 1 2; do
  ICP_NO_SMALL_CTA=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/BIG /"
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/SMALL /"
done
for L in c35_5x5_48_64 c35_3x3_96 c35_3x3_64_96 conv3; do
  ICP_NO_SMALL_CTA=1 timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/BIG /'
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed "s/^/SMALL /"
done
ICP_NO_SMALL_CTA=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 1 --graph --iters 100 | grep SPEED | sed "s/^/BIG /"
timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 1 --graph --iters 100 | grep SPEED | sed "s/^/SMALL /"
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
