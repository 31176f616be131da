cd /root/repo
timeout 600 python -m pytest tests -x -q -m gpu -k "trained_like" 2>&1 | grep -E "assert|Error|passed|failed|rel|AssertionError" | head -20
for i in 1 2; do
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_prev.so timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/prev /"
timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/new  /"
done
for L in conv1 conv2 conv3 c35_1x1_288_256; do
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_prev.so timeout 200 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/prev /"
timeout 200 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/new  /"
done
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
