#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=session26
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/$S.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/$S.log
  tail -n 40 "gpurun_out/$name.log" >> gpurun_out/$S.log
}
run conv_cases 300 python tools/gpu_bringup.py conv
run halo_conv_cases 300 env ICP_HALO_MIN_W=30 ICP_HALO_TALL=1 python tools/gpu_bringup.py conv
run pools 120 python tools/gpu_bringup.py pools
run tower_fp16 300 python tools/gpu_bringup.py tower --precision fp16 --n 2
for L in conv1 conv2 conv3 c35_5x5_48_64 c35_3x3_64_96 c35_3x3_96; do
  run "layer_$L" 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64
done
for L in c35_5x5_48_64 c35_3x3_64_96 c35_3x3_96 conv4; do
  run "layer_halo_$L" 300 env ICP_HALO_MIN_W=30 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64
done
ICP_TIMELINE=gpurun_out/tl_conv1.bin timeout 200 python tools/gpu_bringup.py layer --name conv1 --reps 2 --n 32 > /dev/null 2>&1
ICP_TIMELINE=gpurun_out/tl_conv2.bin timeout 200 python tools/gpu_bringup.py layer --name conv2 --reps 2 --n 32 > /dev/null 2>&1
ICP_HALO_MIN_W=30 ICP_TIMELINE=gpurun_out/tl_c35_5x5.bin timeout 200 python tools/gpu_bringup.py layer --name c35_5x5_48_64 --reps 2 --n 64 > /dev/null 2>&1
run profile_b128 300 python tools/gpu_bringup.py profile --n 128 --precision bf16
run halo_profile_b128 300 env ICP_HALO_MIN_W=30 python tools/gpu_bringup.py profile --n 128 --precision bf16
run speed_b128_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
run speed_b128_halo30 300 env ICP_HALO_MIN_W=30 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
run speed_b1_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 1 --graph --iters 100
run pytest_gpu 900 python -m pytest tests -x -q -m gpu
grep -E "^===.*rc=[1-9]|CASE.*(BAD|EXC)|max \|dprob|SPEED|LAYER|PROFILE|passed|failed|icp:" gpurun_out/$S.log | cut -c1-300
