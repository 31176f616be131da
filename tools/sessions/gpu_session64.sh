#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
ICP_HALO_DIRECT=1 timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"; grep -E "BAD|EXC|icp:|rror" gpurun_out/conv_cases.log | head
ICP_HALO_DIRECT=1 ICP_HALO_MIN_W=30 ICP_HALO_TALL=1 timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases2.log 2>&1; echo "halo-all conv rc=$?"; grep -E "BAD|EXC|icp:|rror" gpurun_out/conv_cases2.log | head
for L in conv1 conv2; do
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 128 | grep LAYER | sed 's/^/STAGED /'
  ICP_HALO_DIRECT=1 timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 128 | grep LAYER | sed "s/^/DIRECT /"
done
for i in 1 2; do
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/STAGED /"
  ICP_HALO_DIRECT=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/DIRECT /"
done
