cd /root/repo
echo "== 2-GPU bench through ShardedTower (NCCL)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 --no-latency 2>&1 | grep -E "^\{|rror|Traceback" | cut -c1-1500
echo "== 2-GPU strong scaling arm"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-latency --strong --sustain-seconds 0 2>&1 | grep -E "^\{|rror|Traceback" | cut -c1-700
echo "== NCCL sharding test + two-engine test"
timeout 900 python -m pytest tests -q -m gpu -k "nccl or two_engines or trained_like" 2>&1 | tail -30
echo "== halo-res 5x5 with unrolled tail"
timeout 120 python tools/gpu_bringup.py layer --name c35_5x5_48_64 --n 128 --precision fp16 --reps 20 2>&1 | grep LAYER | sed "s/^/im2col /"
ICP_HALO_RES=1 timeout 120 python tools/gpu_bringup.py layer --name c35_5x5_48_64 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/res    /"
ICP_HALO_RES=1 timeout 120 python tools/gpu_bringup.py layer --name c35_3x3_96 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/res    /"
ICP_HALO_RES=1 ICP_HALO_RES_MIN_BATCH=1 timeout 300 python tools/gpu_bringup.py conv --only 5x5_same_k48_n64 2>&1 | grep CASE
echo "== timeline of the persistent kernel on a 35x35 1x1"
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so ICP_TIMELINE=gpurun_out/tl_c35_1x1.bin timeout 120 python tools/gpu_bringup.py layer --name c35_1x1_288_256 --n 128 --precision fp16 --reps 2 2>&1 | tail -1
python tools/timeline.py gpurun_out/tl_c35_1x1.bin 5 > gpurun_out/tl_c35_1x1.txt 2>&1
