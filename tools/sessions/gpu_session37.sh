#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tensorflow_web_deploy_b200/_ab/libinception_b200_base.so
timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"; grep -E "BAD|EXC" gpurun_out/conv_cases.log | head
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 | grep dprob
for i in 1 2; do
  ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed 's/^/BASE /'
  for c in 0 16 32 64; do
    ICP_STEM_CHUNK=$c timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/NEW chunk=$c /"
  done
done
for L in conv3 c35_5x5_48_64 c35_3x3_64_96 c35_3x3_96; do
  ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/BASE /'
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/NEW  /'
done
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
