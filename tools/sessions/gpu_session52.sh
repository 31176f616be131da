#!/bin/bash
cd "$(dirname "$0")/../.."
B=tensorflow_web_deploy_b200/_ab/libinception_b200_base.so
for i in 1 2; do
  ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/HEAD /"
  ICP_NO_KGROUP=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/G1 /"
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/G2 /"
done
for L in c35_5x5_48_64 c35_3x3_96 c17_1x7_192 gemm_k1024_n256; do
  ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/HEAD /'
  ICP_NO_KGROUP=1 timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/G1 /'
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed "s/^/G2 /"
done
