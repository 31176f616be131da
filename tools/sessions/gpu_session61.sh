#!/bin/bash
cd "$(dirname "$0")/../.."
D=tensorflow_web_deploy_b200/_ab
for L in conv4 c35_5x5_48_64 c35_1x1_288_256 gemm_k1024_n256 c17_1x7_192; do
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 128 | grep LAYER | sed 's/^/CUR /'
  for v in NOEPI NOTMA_NOEPI NOMMA_NOEPI NONE NONE_NOEPI; do
    ICP_AB_LIB=$D/lib_$v.so timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 128 | grep LAYER | sed "s/^/$v /"
  done
done
