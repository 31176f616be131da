cd /root/repo
for i in 1 2; do
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/lib_base.so timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 --u8 2>&1 | grep SPEED | sed "s/^/base /"
timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 --u8 2>&1 | grep SPEED | sed "s/^/new  /"
done
timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED | sed "s/^/new f32 /"
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/lib_base.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv0 -c 6 python tools/gpu_bringup.py speed --precision fp16 --n 128 --iters 2 --u8 2>&1 | grep -E "conv0_tc|gpu__time" | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv0 -c 6 python tools/gpu_bringup.py speed --precision fp16 --n 128 --iters 2 --u8 2>&1 | grep -E "conv0_tc|gpu__time" | tail -4
