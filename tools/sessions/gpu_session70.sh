#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tensorflow_web_deploy_b200/_ab/libinception_b200_base.so
timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"; grep -E "BAD|EXC|icp:|rror" gpurun_out/conv_cases.log | head
for i in 1 2; do
  ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/HEAD /"
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/SPEC /"
done
for L in conv4 c35_5x5_48_64 c17_1x7_192 c35_1x1_288_256 c17_1x1_768_704; do
  ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 128 | grep LAYER | sed 's/^/HEAD /'
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 128 | grep LAYER | sed "s/^/SPEC /"
done
