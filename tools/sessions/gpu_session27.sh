#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
ICP_TIMELINE=gpurun_out/tl_conv1.bin timeout 200 python tools/gpu_bringup.py layer --name conv1 --reps 2 --n 32 > gpurun_out/tl1.log 2>&1
ICP_TIMELINE=gpurun_out/tl_conv2.bin timeout 200 python tools/gpu_bringup.py layer --name conv2 --reps 2 --n 32 > gpurun_out/tl2.log 2>&1
tail -2 gpurun_out/tl1.log gpurun_out/tl2.log
