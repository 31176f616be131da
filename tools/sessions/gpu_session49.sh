#!/bin/bash
cd "$(dirname "$0")/../.."
D=tensorflow_web_deploy_b200/_ab
export ICP_NO_PAIR_BOX=1
for L in c35_5x5_48_64 c17_1x7_192 gemm_k1024_n256 conv4 c35_1x1_288_256; do
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/CUR /'
  for v in SPIN SPIN_NONE; do
    ICP_AB_LIB=$D/lib_$v.so timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed "s/^/$v /"
  done
done
for i in 1 2; do
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/CUR /"
  ICP_AB_LIB=$D/lib_SPIN.so timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/SPIN /"
done
