#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
D=tensorflow_web_deploy_b200/_ab
ICP_AB_LIB=$D/lib_G4.so timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"; grep -E "BAD|EXC" gpurun_out/conv_cases.log | head
for i in 1 2; do
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/CUR /"
  ICP_AB_LIB=$D/lib_G4.so timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/G4 /"
done
for L in c35_5x5_48_64 c35_3x3_96 c17_1x7_192 gemm_k1024_n256 conv4; do
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/CUR /'
  ICP_AB_LIB=$D/lib_G4.so timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed "s/^/G4 /"
done
