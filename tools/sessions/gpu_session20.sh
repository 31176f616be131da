#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session20.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/session20.log
  tail -n 40 "gpurun_out/$name.log" >> gpurun_out/session20.log
}
for L in gemm_k1024_n256 gemm_k1024_n192 gemm_k1024_n128 conv3x3_k256_n256; do
  run "base_$L" 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64
  run "split_$L" 300 env ICP_SPLIT_N=1 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64
done
run split_conv_cases 300 env ICP_SPLIT_N=1 python tools/gpu_bringup.py conv
run split_speed_b128_graph 300 env ICP_SPLIT_N=1 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
run speed_b128_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
grep -E "^===.*rc|CASE.*(BAD|EXC)|SPEED|LAYER|icp:" gpurun_out/session20.log | cut -c1-300
