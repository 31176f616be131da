cd /root/repo
echo "== pair correctness"
ICP_PAIR=1 ICP_PAIR_MIN_BATCH=1 timeout 600 python tools/gpu_bringup.py conv 2>&1 | grep -E "CASE|bring-up|bad rows" | grep -v " ok " | head -30
ICP_PAIR=1 ICP_PAIR_MIN_BATCH=1 timeout 600 python tools/gpu_bringup.py conv 2>&1 | grep -c " ok "
echo "== pair timing"
for L in c35_3x3_96 c35_3x3_64_96 c35_5x5_48_64 conv4; do
  timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/im2col /"
  ICP_PAIR=1 timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/pair   /"
done
for L in c35_3x3_96 c35_5x5_48_64; do
ICP_PAIR=1 ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so ICP_TIMELINE=gpurun_out/tlp_$L.bin timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 2 2>&1 | tail -1
python tools/timeline.py gpurun_out/tlp_$L.bin > gpurun_out/tlp_$L.txt 2>&1
done
echo "== tower with pair"
ICP_PAIR=1 timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 9 2>&1 | tail -4
ICP_PAIR=1 timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph 2>&1 | grep SPEED
timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph 2>&1 | grep SPEED
