#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
ICP_TIMELINE=gpurun_out/tl_gemm_k1024_n256.bin timeout 200 python tools/gpu_bringup.py layer --name gemm_k1024_n256 --reps 2 --n 64 > gpurun_out/tl1.log 2>&1
ICP_TIMELINE=gpurun_out/tl_c17_1x1.bin timeout 200 python tools/gpu_bringup.py layer --name c17_1x1_768_704 --reps 2 --n 64 > gpurun_out/tl2.log 2>&1
ICP_TIMELINE=gpurun_out/tl_c35_1x1.bin timeout 200 python tools/gpu_bringup.py layer --name c35_1x1_288_256 --reps 10 --n 64 > gpurun_out/tl3.log 2>&1
cat gpurun_out/tl1.log gpurun_out/tl2.log gpurun_out/tl3.log | grep LAYER
