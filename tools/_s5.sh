cd /root/repo
python -m pytest tests -q -m gpu -k "jpeg or micro_batch" -s 2>&1 | grep -E "JPEG |passed|failed|Error|assert " | head -40
for L in c35_3x3_96 c35_3x3_64_96 c35_5x5_48_64 conv4 c17_1x7_192; do
  python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/im2col /"
  ICP_HALO_MIN_W=30 ICP_HALO_TALL=1 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/halo   /"
done
for L in c35_3x3_96 c35_5x5_48_64; do
  ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so ICP_TIMELINE=gpurun_out/tl_$L.bin ICP_HALO_MIN_W=30 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 2 2>&1 | grep LAYER
  python tools/timeline.py gpurun_out/tl_$L.bin > gpurun_out/tl_$L.txt 2>&1
done
