#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
D=tensorflow_web_deploy_b200/_ab
export ICP_NO_PAIR_BOX=1
ICP_AB_LIB=$D/lib_TLNONE.so ICP_TIMELINE=gpurun_out/tl_none_conv4.bin timeout 200 python tools/gpu_bringup.py layer --name conv4 --reps 5 --n 64 | grep LAYER
ICP_AB_LIB=$D/libinception_b200_tl.so ICP_TIMELINE=gpurun_out/tl_conv4.bin timeout 200 python tools/gpu_bringup.py layer --name conv4 --reps 5 --n 64 | grep LAYER
