#!/bin/bash
# Per-op profile + ncu launch list + ncu full capture of the first conv_igemm launches.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session2.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/session2.log
  tail -n 130 "gpurun_out/$name.log" >> gpurun_out/session2.log
}
run profile_b128 300 python tools/gpu_bringup.py profile --n 128 --precision bf16
run profile_b1 300 python tools/gpu_bringup.py profile --n 1 --precision bf16 --reps 20
run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_b128.csv python tools/gpu_bringup.py speed --n 128 --iters 0
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 93 -c 30 -o gpurun_out/conv_igemm_r01 python tools/gpu_bringup.py speed --n 128 --iters 0
ls -la gpurun_out
grep -E "^===|PROFILE|OP " gpurun_out/session2.log | tail -250
