#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session7.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/session7.log
  tail -n 60 "gpurun_out/$name.log" >> gpurun_out/session7.log
}
run profile_b128 300 python tools/gpu_bringup.py profile --n 128 --precision bf16
run speed_b128_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
run pytest_gpu 900 python -m pytest tests -x -q -m gpu
run ncu_halo 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 2 -o gpurun_out/halo_r01 python tools/gpu_bringup.py speed --n 128 --iters 1
run ncu_persist 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_persistent -s 30 -c 12 -o gpurun_out/persist_r01 python tools/gpu_bringup.py speed --n 128 --iters 1
grep -E "^===|PROFILE|SPEED|passed|failed" gpurun_out/session7.log | tail -30
grep "^OP " gpurun_out/profile_b128.log | head -8
