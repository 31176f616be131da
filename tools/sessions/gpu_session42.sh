#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so
ICP_AB_LIB=$T ICP_TIMELINE=gpurun_out/tl_conv4.bin timeout 200 python tools/gpu_bringup.py layer --name conv4 --reps 5 --n 64 | grep LAYER
ICP_AB_LIB=$T ICP_TIMELINE=gpurun_out/tl_c17_1x7.bin timeout 200 python tools/gpu_bringup.py layer --name c17_1x7_192 --reps 5 --n 64 | grep LAYER
