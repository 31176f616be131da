#!/bin/bash
# First GPU bring-up: every step under its own timeout, logs into gpurun_out/.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session1.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "=== $name rc=$rc" | tee -a gpurun_out/session1.log
  tail -n 60 "gpurun_out/$name.log" >> gpurun_out/session1.log
}
run conv_default 300 python tools/gpu_bringup.py conv
ICP_FORCE_IM2COL=1 run conv_im2col1x1 200 env ICP_FORCE_IM2COL=1 python tools/gpu_bringup.py conv --only 1x1_k192_n64
run pools 120 python tools/gpu_bringup.py pools
run tower_fp16 400 python tools/gpu_bringup.py tower --precision fp16 --n 2
run tower_bf16 400 python tools/gpu_bringup.py tower --precision bf16 --n 2
run tower_fp16_graph 400 python tools/gpu_bringup.py tower --precision fp16 --n 2 --graph
run speed_b128 300 python tools/gpu_bringup.py speed --precision bf16 --n 128
run speed_b128_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
run speed_b1_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 1 --graph --iters 50
cat gpurun_out/session1.log | grep -E "^===|CASE|EP |SPEED|POOLS|max \|dprob" | tail -150
