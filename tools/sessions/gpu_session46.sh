#!/bin/bash
cd "$(dirname "$0")/../.."
D=tensorflow_web_deploy_b200/_ab
for L in c35_5x5_48_64 c35_3x3_96 c17_1x7_192 gemm_k1024_n256 conv4 c17_1x1_768_704; do
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/CUR /'
  ICP_AB_LIB=$D/lib_DUPW.so timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed "s/^/DUPW /"
done
