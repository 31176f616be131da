cd /root/repo
echo "== dual-tile correctness (forced everywhere)"
ICP_DUAL=1 ICP_DUAL_MIN=0 ICP_DUAL_MAX=100000 ICP_DUAL_MIN_TAPS=1 timeout 600 python tools/gpu_bringup.py conv 2>&1 | grep -E "CASE|bring-up|bad rows" | grep -v " ok " | head -20
ICP_DUAL=1 ICP_DUAL_MIN=0 ICP_DUAL_MAX=100000 ICP_DUAL_MIN_TAPS=1 timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 5 2>&1 | tail -2
echo "== timing"
for L in c17_1x7_192 c17_1x1_768_704 c35_3x3_96 c8_3x3_448 conv4; do
  timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/base /"
  ICP_DUAL=1 ICP_DUAL_MIN=0 ICP_DUAL_MAX=100000 ICP_DUAL_MIN_TAPS=1 timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/dual /"
done
for i in 1 2; do
timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/base        /"
ICP_DUAL=1 timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/dual 1.5-3  /"
ICP_DUAL=1 ICP_DUAL_MAX=10 timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/dual 1.5-10 /"
ICP_DUAL=1 ICP_DUAL_MAX=10 ICP_DUAL_MIN_TAPS=1 timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/dual +1x1   /"
done
