cd /root/repo
L=c35_3x3_64_96
python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/im2col /"
ICP_HALO_MIN_W=30 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/halo-stream /"
ICP_HALO_MIN_W=30 ICP_HALO_PREFER_RESIDENT=1 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/halo-resident /"
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so ICP_TIMELINE=gpurun_out/tlr_$L.bin ICP_HALO_MIN_W=30 ICP_HALO_PREFER_RESIDENT=1 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 2 2>&1 | tail -2
python tools/timeline.py gpurun_out/tlr_$L.bin > gpurun_out/tlr_$L.txt 2>&1
ICP_HALO_MIN_W=30 ICP_HALO_PREFER_RESIDENT=1 python tools/gpu_bringup.py conv --only halo_35_k96_3x3_n96 2>&1 | tail -3
python tools/gpu_bringup.py layer --name conv2 --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so ICP_TIMELINE=gpurun_out/tl_conv2.bin python tools/gpu_bringup.py layer --name conv2 --n 128 --precision fp16 --reps 2 2>&1 | tail -1
python tools/timeline.py gpurun_out/tl_conv2.bin > gpurun_out/tl_conv2.txt 2>&1
