#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"; grep -E "BAD|EXC" gpurun_out/conv_cases.log | head
ICP_HALO_MIN_W=30 ICP_HALO_TALL=1 timeout 300 python tools/gpu_bringup.py conv > gpurun_out/halo_conv_cases.log 2>&1; echo "halo conv rc=$?"; grep -E "BAD|EXC" gpurun_out/halo_conv_cases.log | head
for i in 1 2; do
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/NEW /"
  ICP_HALO_MIN_W=30 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/NEW halo30 /"
  ICP_HALO_MIN_W=30 ICP_HALO_TALL=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/NEW halo30tall /"
done
timeout 300 python tools/gpu_bringup.py profile --n 128 --precision bf16 > gpurun_out/profile_b128.log 2>&1; grep PROFILE gpurun_out/profile_b128.log
ICP_HALO_MIN_W=30 ICP_HALO_TALL=1 timeout 300 python tools/gpu_bringup.py profile --n 128 --precision bf16 > gpurun_out/profile_b128_halo.log 2>&1; grep PROFILE gpurun_out/profile_b128_halo.log
timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 1 --graph --iters 100 | grep SPEED
timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 8 --graph --iters 50 | grep SPEED
timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 32 --graph --iters 30 | grep SPEED
