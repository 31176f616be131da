#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so
for d in 0 1; do
ICP_HALO_DIRECT=$d ICP_AB_LIB=$T ICP_TIMELINE=gpurun_out/tl_conv2_d$d.bin timeout 200 python tools/gpu_bringup.py layer --name conv2 --reps 3 --n 32 | grep LAYER
ICP_HALO_DIRECT=$d ICP_AB_LIB=$T ICP_TIMELINE=gpurun_out/tl_conv1_d$d.bin timeout 200 python tools/gpu_bringup.py layer --name conv1 --reps 3 --n 32 | grep LAYER
done
