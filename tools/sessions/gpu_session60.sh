#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
D=tensorflow_web_deploy_b200/_ab
for L in conv4 c35_5x5_48_64 gemm_k1024_n256; do
ICP_AB_LIB=$D/lib_TLCL.so ICP_TIMELINE=gpurun_out/tl_cl_$L.bin timeout 200 python tools/gpu_bringup.py layer --name $L --reps 3 --n 64 | grep LAYER
done
