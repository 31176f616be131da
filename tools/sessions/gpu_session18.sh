#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session18.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/session18.log
  tail -n 40 "gpurun_out/$name.log" >> gpurun_out/session18.log
}
run tower_fp16_graph 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 --graph
run tower_fp16 300 python tools/gpu_bringup.py tower --precision fp16 --n 2
run speed_b1_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 1 --graph --iters 100
run speed_b1_graph_nolanes 300 env ICP_LANES_MAX_BATCH=0 python tools/gpu_bringup.py speed --precision bf16 --n 1 --graph --iters 100
run speed_b8_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 8 --graph --iters 50
run speed_b8_graph_nolanes 300 env ICP_LANES_MAX_BATCH=0 python tools/gpu_bringup.py speed --precision bf16 --n 8 --graph --iters 50
run speed_b32_lanes 300 env ICP_LANES_MAX_BATCH=128 python tools/gpu_bringup.py speed --precision bf16 --n 32 --graph --iters 30
run speed_b32_nolanes 300 env ICP_LANES_MAX_BATCH=0 python tools/gpu_bringup.py speed --precision bf16 --n 32 --graph --iters 30
run speed_b128_lanes 300 env ICP_LANES_MAX_BATCH=128 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
run speed_b128_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
for L in gemm_k1024_n256 gemm_k1024_n192 gemm_k1024_n128 conv3x3_k256_n256; do run "layer_$L" 300 python tools/gpu_bringup.py layer --name $L --reps 10 --n 64; done
run pytest_gpu 900 python -m pytest tests -x -q -m gpu
grep -E "^===|max \|dprob|SPEED|LAYER|passed|failed|icp:" gpurun_out/session18.log | cut -c1-300
