#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session12.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/session12.log
  tail -n 12 "gpurun_out/$name.log" >> gpurun_out/session12.log
}
for L in conv4 c17_1x7_192 c17_1x1_768_704 c35_3x3_96 c8_3x3_448; do
  run "t_halo_$L" 200 python tools/gpu_bringup.py layer --name $L --reps 20
  run "t_nohalo_$L" 200 env ICP_NO_HALO=1 python tools/gpu_bringup.py layer --name $L --reps 20
done
NCU="ncu --set full --clock-control none --import-source on"
run ncu_halo_conv4 600 $NCU -k regex:conv_halo -s 3 -c 1 -o gpurun_out/halo_conv4 python tools/gpu_bringup.py layer --name conv4 --reps 2
run ncu_persist_conv4 600 env ICP_NO_HALO=1 $NCU -k regex:conv_igemm_persistent -s 3 -c 1 -o gpurun_out/persist_conv4 python tools/gpu_bringup.py layer --name conv4 --reps 2
run ncu_persist_1x7 600 env ICP_NO_HALO=1 $NCU -k regex:conv_igemm_persistent -s 3 -c 1 -o gpurun_out/persist_1x7 python tools/gpu_bringup.py layer --name c17_1x7_192 --reps 2
run ncu_halo_1x7 600 $NCU -k regex:conv_halo -s 3 -c 1 -o gpurun_out/halo_1x7 python tools/gpu_bringup.py layer --name c17_1x7_192 --reps 2
grep -E "^===|LAYER" gpurun_out/session12.log
ls -la gpurun_out/*.ncu-rep
