#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 2>&1 | grep -E "dprob|rror|icp|EP conv2|EP pool1" | head -6
for i in 1 2; do
  ICP_NO_HALO_TMA=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/DENSE /"
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/PITCHED /"
done
timeout 300 python tools/gpu_bringup.py profile --n 128 --precision bf16 > gpurun_out/profile_b128.log 2>&1; grep -E "PROFILE|^OP conv2 |^OP conv1 |^OP max_pool" gpurun_out/profile_b128.log | cut -c1-140 | head -5
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
