#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
B=tensorflow_web_deploy_b200/_ab/libinception_b200_base.so
T=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so
timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"; grep -E "BAD|EXC" gpurun_out/conv_cases.log | head
for i in 1 2; do
  ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed 's/^/BASE /'
  timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph | grep SPEED | sed "s/^/NEW /"
done
for L in conv4 c35_5x5_48_64 c35_3x3_96 gemm_k1024_n256 c17_1x1_768_704 c35_1x1_288_256 c17_1x7_192; do
  ICP_AB_LIB=$B timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/BASE /'
  timeout 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64 | grep LAYER | sed 's/^/NEW  /'
done
ICP_AB_LIB=$T ICP_TIMELINE=gpurun_out/tl_c35_5x5.bin timeout 200 python tools/gpu_bringup.py layer --name c35_5x5_48_64 --reps 5 --n 64 > /dev/null 2>&1
ICP_AB_LIB=$T ICP_TIMELINE=gpurun_out/tl_gemm.bin timeout 200 python tools/gpu_bringup.py layer --name gemm_k1024_n256 --reps 5 --n 64 > /dev/null 2>&1
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
