#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python tools/gpu_bringup.py layer --name conv4 --reps 3 --n 8 2>&1 | tail -2
timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"; grep -E "BAD|EXC|icp:|rror" gpurun_out/conv_cases.log | head
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 2>&1 | grep -E "dprob|rror|icp" | head -3
for i in 1 2; do
  ICP_NO_CLUSTER=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/SINGLE /"
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/CLUSTER /"
done
for L in conv4 c35_5x5_48_64 c35_3x3_96 c17_1x7_192 c35_1x1_288_256 gemm_k1024_n256; do
  ICP_NO_CLUSTER=1 timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 128 | grep LAYER | sed 's/^/SINGLE /'
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 128 | grep LAYER | sed "s/^/CLUSTER /"
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
