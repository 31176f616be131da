#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
ICP_NO_DUAL_M=1 timeout 300 python tools/gpu_bringup.py profile --n 128 --precision bf16 > gpurun_out/profile_single.log 2>&1; grep PROFILE gpurun_out/profile_single.log
timeout 300 python tools/gpu_bringup.py profile --n 128 --precision bf16 > gpurun_out/profile_dual.log 2>&1; grep PROFILE gpurun_out/profile_dual.log
