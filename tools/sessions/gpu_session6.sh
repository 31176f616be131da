#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session6.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/session6.log
  tail -n 60 "gpurun_out/$name.log" >> gpurun_out/session6.log
}
run conv_cases 300 python tools/gpu_bringup.py conv
run tower_fp16 300 python tools/gpu_bringup.py tower --precision fp16 --n 2
run profile_b128 300 python tools/gpu_bringup.py profile --n 128 --precision bf16
cp gpurun_out/profile_ops_n128_bf16.json gpurun_out/profile_ops_n128_bf16_halo.json
run nohalo_profile_b128 300 env ICP_NO_HALO=1 python tools/gpu_bringup.py profile --n 128 --precision bf16
run speed_b128_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
run pytest_gpu 900 python -m pytest tests -x -q -m gpu
grep -E "^===|CASE.*(BAD|EXC)|max \|dprob|PROFILE|SPEED|passed|failed|^\{" gpurun_out/session6.log | cut -c1-600 | tail -40
grep "^OP " gpurun_out/profile_b128.log | head -24
