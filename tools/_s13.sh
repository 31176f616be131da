cd /root/repo
echo "== resident N-split halo: correctness"
ICP_HALO_RES=1 ICP_HALO_RES_MIN_BATCH=1 timeout 600 python tools/gpu_bringup.py conv 2>&1 | grep -E "CASE|bring-up|bad rows" | grep -v " ok " | head -20
echo "== timing"
for L in c35_3x3_96 c35_3x3_64_96 c35_5x5_48_64; do
  timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/im2col /"
  ICP_HALO_RES=1 timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/res    /"
  ICP_HALO_RES=1 ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so ICP_TIMELINE=gpurun_out/tlr_$L.bin timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 2 2>&1 | tail -1
  python tools/timeline.py gpurun_out/tlr_$L.bin > gpurun_out/tlr_$L.txt 2>&1
done
ICP_HALO_RES=1 ICP_HALO_RES_MAX_W=100 timeout 120 python tools/gpu_bringup.py layer --name conv4 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/res conv4 /"
echo "== tower"
ICP_HALO_RES=1 timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 9 2>&1 | tail -2
for i in 1 2; do
ICP_HALO_RES=1 timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/res  /"
timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/base /"
done
timeout 600 python -m pytest tests -x -q -m gpu -k "trained_like" 2>&1 | tail -3
