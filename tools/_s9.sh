cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_prev.so timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/prev /"
timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/new  /"
done
timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 1 --graph --iters 200 2>&1 | grep SPEED | sed "s/^/new  /"
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_prev.so timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 1 --graph --iters 200 2>&1 | grep SPEED | sed "s/^/prev /"
timeout 200 python tools/gpu_bringup.py profile --n 128 --precision fp16 2>&1 | grep -E "conv3|fused 1x1|PROFILE" | cut -c1-150
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_prev.so timeout 200 python tools/gpu_bringup.py profile --n 128 --precision fp16 2>&1 | grep -E "conv3|fused 1x1|PROFILE" | cut -c1-150 | sed "s/^/prev /"
ICP_PAIR=1 ICP_PAIR_MIN_BATCH=1 timeout 200 python tools/gpu_bringup.py conv --only 3x3_same_k32_n64 2>&1 | grep CASE
ICP_PAIR=1 ICP_PAIR_MIN_BATCH=1 timeout 200 python tools/gpu_bringup.py conv --only 3x3_valid_k32_n32 2>&1 | grep CASE
