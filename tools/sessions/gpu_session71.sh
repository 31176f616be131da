#!/bin/bash
cd "$(dirname "$0")/../.."
B=tensorflow_web_deploy_b200/_ab/libinception_b200_base.so
for L in conv1 conv2; do
  ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 128 | grep LAYER | sed 's/^/HEAD /'
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 128 | grep LAYER | sed "s/^/STEM /"
done
ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/HEAD /"
timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/STEM /"
