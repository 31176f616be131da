cd /root/repo
for L in c35_3x3_96 c35_5x5_48_64; do
  ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so ICP_TIMELINE=gpurun_out/tl_$L.bin ICP_HALO_MIN_W=30 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 2 2>&1 | tail -3
  python tools/timeline.py gpurun_out/tl_$L.bin > gpurun_out/tl_$L.txt 2>&1
  ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so ICP_TIMELINE=gpurun_out/tli_$L.bin python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 2 2>&1 | tail -3
  python tools/timeline.py gpurun_out/tli_$L.bin > gpurun_out/tli_$L.txt 2>&1
done
