#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session4.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/session4.log
  tail -n 60 "gpurun_out/$name.log" >> gpurun_out/session4.log
}
run conv_cases 300 python tools/gpu_bringup.py conv
run tower_fp16 300 python tools/gpu_bringup.py tower --precision fp16 --n 2
run tower_bf16 300 python tools/gpu_bringup.py tower --precision bf16 --n 2
run profile_b128 300 python tools/gpu_bringup.py profile --n 128 --precision bf16
run nohalo_profile_b128 300 env ICP_NO_HALO=1 python tools/gpu_bringup.py profile --n 128 --precision bf16
run pytest_gpu 900 python -m pytest tests -x -q -m gpu
run bench1 600 python bench.py --gpus 1 --steps 20 --warmup 3
run bench_ref 400 python bench.py --impl reference --steps 5 --warmup 3
grep -E "^===|CASE.*(BAD|EXC)|max \|dprob|PROFILE|passed|failed|^\{" gpurun_out/session4.log | cut -c1-600 | tail -60
