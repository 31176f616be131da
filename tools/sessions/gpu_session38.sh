#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tensorflow_web_deploy_b200/_ab/libinception_b200_base.so
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 | grep dprob
timeout 300 python tools/gpu_bringup.py tower --precision bf16 --n 3 | grep dprob
for i in 1 2; do
  ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed 's/^/BASE /'
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/NEW /"
  ICP_CONV_V2_ALL=1 timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/NEW v2all /"
done
timeout 300 python tools/gpu_bringup.py profile --n 128 --precision bf16 > gpurun_out/profile_b128.log 2>&1; grep PROFILE gpurun_out/profile_b128.log
ICP_CONV_V2_ALL=1 timeout 300 python tools/gpu_bringup.py profile --n 128 --precision bf16 > gpurun_out/profile_b128_v2all.log 2>&1; grep PROFILE gpurun_out/profile_b128_v2all.log
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
