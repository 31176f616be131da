#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so
for L in c35_5x5_48_64 c35_3x3_96 c35_3x3_64_96; do
  ICP_AB_LIB=$T ICP_HALO_MIN_W=30 ICP_TIMELINE=gpurun_out/tl_halo_$L.bin timeout 200 python tools/gpu_bringup.py layer --name $L --reps 5 --n 64 | grep LAYER | sed "s/^/HALO-tl /"
  ICP_AB_LIB=$T ICP_TIMELINE=gpurun_out/tl_im2col_$L.bin timeout 200 python tools/gpu_bringup.py layer --name $L --reps 5 --n 64 | grep LAYER | sed "s/^/IM2COL-tl /"
  ICP_HALO_MIN_W=30 timeout 200 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed "s/^/HALO /"
  timeout 200 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed "s/^/IM2COL /"
done
L=c17_1x7_192
ICP_AB_LIB=$T ICP_HALO_MIN_W=30 ICP_HALO_TALL=1 ICP_TIMELINE=gpurun_out/tl_halo_$L.bin timeout 200 python tools/gpu_bringup.py layer --name $L --reps 5 --n 64 | grep LAYER | sed "s/^/HALO-tl /"
ICP_AB_LIB=$T ICP_TIMELINE=gpurun_out/tl_im2col_$L.bin timeout 200 python tools/gpu_bringup.py layer --name $L --reps 5 --n 64 | grep LAYER | sed "s/^/IM2COL-tl /"
ICP_HALO_MIN_W=30 ICP_HALO_TALL=1 timeout 200 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed "s/^/HALO /"
timeout 200 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed "s/^/IM2COL /"
