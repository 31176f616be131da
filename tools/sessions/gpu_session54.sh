#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 2>&1 | grep -E "dprob|rror|icp" | head -3
timeout 300 python tools/gpu_bringup.py tower --precision bf16 --n 3 2>&1 | grep -E "dprob|rror|icp" | head -3
for i in 1 2; do
  ICP_CONV0_FMA=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/FMA /"
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/TC /"
done
timeout 300 python tools/gpu_bringup.py profile --n 128 --precision bf16 > gpurun_out/profile_b128.log 2>&1; grep -E "PROFILE|^OP conv0|^OP conv1 " gpurun_out/profile_b128.log | cut -c1-160
timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 1 --graph --iters 100 | grep SPEED
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
