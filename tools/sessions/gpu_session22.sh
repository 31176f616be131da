#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for L in gemm_k1024_n256 gemm_k1024_n128 c17_1x7_192 conv4; do
  ICP_TIMELINE=gpurun_out/tl_$L.bin timeout 200 python tools/gpu_bringup.py layer --name $L --reps 2 --n 64 > gpurun_out/tl_$L.log 2>&1
  tail -1 gpurun_out/tl_$L.log
done
ls -la gpurun_out
