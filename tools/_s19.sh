cd /root/repo
mkdir -p gpurun_out/r02
TL=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so
for L in conv1 conv2; do
  ICP_AB_LIB=$TL ICP_TIMELINE=gpurun_out/r02/tl_$L.bin timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 2 > /dev/null 2>&1
  python tools/timeline.py gpurun_out/r02/tl_$L.bin > gpurun_out/r02/tl_$L.txt 2>&1; head -30 gpurun_out/r02/tl_$L.txt
done
echo "== conv1 T=8"
ICP_HALO_T=8 timeout 120 python tools/gpu_bringup.py layer --name conv1 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror"
echo "== conv4 halo"
timeout 120 python tools/gpu_bringup.py layer --name conv4 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror"
ICP_HALO_MIN_W=70 timeout 120 python tools/gpu_bringup.py layer --name conv4 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror"
echo "== tower halo>=30"
ICP_HALO_MIN_W=30 timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED
ICP_HALO_MIN_W=70 timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED
ICP_HALO_T=8 timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED
