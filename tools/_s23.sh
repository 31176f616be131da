cd /root/repo
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 5 2>&1 | tail -1
timeout 300 python tools/gpu_bringup.py tower --precision bf16 --n 3 2>&1 | tail -1
timeout 600 python tools/gpu_bringup.py conv 2>&1 | grep -E "bring-up"
for L in conv1 conv2; do
  ICP_HALO_MAX_STRIPS=2 timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/2 strips /"
  timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/default  /"
done
ICP_HALO_DIRECT=1 timeout 120 python tools/gpu_bringup.py layer --name conv2 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/direct   /"
ICP_HALO_DIRECT=1 ICP_HALO_T=2 timeout 120 python tools/gpu_bringup.py layer --name conv2 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/direct T2 /"
ICP_HALO_T=2 timeout 120 python tools/gpu_bringup.py layer --name conv1 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/conv1 T2 /"
for i in 1 2; do
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/lib_base.so timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 --u8 2>&1 | grep SPEED | sed "s/^/base /"
timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 --u8 2>&1 | grep SPEED | sed "s/^/new  /"
done
timeout 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph --iters 20 --u8 2>&1 | grep SPEED | sed "s/^/new bf16 /"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv0 -c 6 python tools/gpu_bringup.py speed --precision fp16 --n 128 --iters 2 --u8 2>&1 | grep -E "gpu__time" | tail -2
timeout 900 python -m pytest tests -x -q -m gpu -k "parity or batch or ragged or u8 or rgb" 2>&1 | tail -3
