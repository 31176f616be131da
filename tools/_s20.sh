cd /root/repo
mkdir -p gpurun_out/r02
TL=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so
echo "== correctness (box rows 2)"
ICP_HALO_BOX_ROWS=2 timeout 600 python tools/gpu_bringup.py conv 2>&1 | grep -E "CASE|bring-up|bad rows" | grep -v " ok " | head -20
ICP_HALO_BOX_ROWS=2 timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 5 2>&1 | tail -1
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 5 2>&1 | tail -1
echo "== layers"
for L in conv1 conv2; do
  timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/box=all /"
  for B in 2 4; do
  ICP_HALO_BOX_ROWS=$B timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/box=$B   /"
  done
  ICP_HALO_BOX_ROWS=2 ICP_AB_LIB=$TL ICP_TIMELINE=gpurun_out/r02/tl_$L.bin timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 2 > /dev/null 2>&1
  python tools/timeline.py gpurun_out/r02/tl_$L.bin > gpurun_out/r02/tl_${L}_box2.txt 2>&1; sed -n 1,12p gpurun_out/r02/tl_${L}_box2.txt; sed -n 40,50p gpurun_out/r02/tl_${L}_box2.txt
done
ICP_HALO_BOX_ROWS=2 ICP_HALO_T=8 timeout 120 python tools/gpu_bringup.py layer --name conv1 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/box=2 T=8 /"
echo "== conv0 (two epilogue teams)"
timeout 300 python tools/gpu_bringup.py profile --precision fp16 --n 128 2>&1 | grep -E "conv0|conv1|conv2" | head -3
for i in 1 2; do
timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/box=all /"
ICP_HALO_BOX_ROWS=2 timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/box=2   /"
done
