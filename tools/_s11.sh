cd /root/repo
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 --calibrated 2>&1 | grep -E "^EP|max \|dprob" | cut -c1-120
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 2 --calibrated --fc-scale 1 2>&1 | grep -E "max \|dprob"
timeout 600 python -m pytest tests -x -q -m gpu -k "fusion or ragged or full_size or tower_matches" 2>&1 | tail -5
for i in 1 2; do
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_prev.so timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/prev /"
timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/new  /"
ICP_FUSE_POOL_MIN_BATCH=100000 timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/nofuse /"
done
for L in conv1 conv2; do
timeout 200 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/new  /"
done
