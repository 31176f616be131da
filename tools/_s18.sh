cd /root/repo
mkdir -p gpurun_out/r02
echo "== correctness"
timeout 600 python tools/gpu_bringup.py conv 2>&1 | grep -E "CASE|bring-up|bad rows" | grep -v " ok " | head -20
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 5 2>&1 | tail -2
timeout 300 python tools/gpu_bringup.py tower --precision bf16 --n 3 2>&1 | tail -1
echo "== layers"
for L in conv1 conv2; do
  ICP_HALO_CP=0 timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/tma /"
  timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/cp  /"
done
for i in 1 2; do
ICP_HALO_CP=0 timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/tma /"
timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/cp  /"
done
timeout 300 python tools/gpu_bringup.py profile --precision fp16 --n 128 2>&1 | grep -E "conv0|conv1|conv2|conv3|conv4|max_pool" | head -8
timeout 900 python -m pytest tests -x -q -m gpu -k "parity or smoke or sample" 2>&1 | tail -3
