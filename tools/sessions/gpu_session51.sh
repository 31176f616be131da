#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"; grep -E "BAD|EXC|icp:|rror" gpurun_out/conv_cases.log | head
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 2>&1 | grep -E "dprob|rror|icp" | head -3
for i in 1 2; do
  ICP_NO_KGROUP=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/G1 /"
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/G2 /"
done
for L in c35_5x5_48_64 c35_3x3_96 c35_3x3_64_96 conv3 gemm_k1024_n128; do
  ICP_NO_KGROUP=1 timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/G1 /'
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed "s/^/G2 /"
done
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
