cd /root/repo
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5
for i in 1 2; do
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_prev.so timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/prev /"
timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/new  /"
ICP_FUSE_POOL_MIN_BATCH=100000 timeout 200 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/nofuse /"
done
timeout 200 python tools/gpu_bringup.py layer --name conv1 --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so ICP_TIMELINE=gpurun_out/tl_conv1.bin timeout 100 python tools/gpu_bringup.py layer --name conv1 --n 128 --precision fp16 --reps 2 2>&1 | tail -1
python tools/timeline.py gpurun_out/tl_conv1.bin > gpurun_out/tl_conv1.txt 2>&1
