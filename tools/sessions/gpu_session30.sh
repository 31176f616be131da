#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=session30
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/$S.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/$S.log
  tail -n 40 "gpurun_out/$name.log" >> gpurun_out/$S.log
}
run conv_cases 300 python tools/gpu_bringup.py conv
run halo_conv_cases 300 env ICP_HALO_MIN_W=30 ICP_HALO_TALL=1 python tools/gpu_bringup.py conv
for O in 0 1; do
for L in conv1 conv2 c35_1x1_288_256 c17_1x1_768_704; do
  run "layer_o${O}_$L" 300 env ICP_EPI_ORDER=$O python tools/gpu_bringup.py layer --name $L --reps 10 --n 64
done
ICP_EPI_ORDER=$O ICP_TIMELINE=gpurun_out/tl_o${O}_conv2.bin timeout 200 python tools/gpu_bringup.py layer --name conv2 --reps 2 --n 32 > /dev/null 2>&1
ICP_EPI_ORDER=$O ICP_TIMELINE=gpurun_out/tl_o${O}_c35_1x1.bin timeout 200 python tools/gpu_bringup.py layer --name c35_1x1_288_256 --reps 2 --n 64 > /dev/null 2>&1
run speed_b128_o$O 300 env ICP_EPI_ORDER=$O python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
done
run tower_fp16 300 python tools/gpu_bringup.py tower --precision fp16 --n 2
run profile_b128 300 python tools/gpu_bringup.py profile --n 128 --precision bf16
grep -E "^===.*rc=[1-9]|CASE.*(BAD|EXC)|max \|dprob|SPEED|LAYER|PROFILE|passed|failed|icp:" gpurun_out/$S.log | cut -c1-300
