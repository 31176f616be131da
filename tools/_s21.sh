cd /root/repo
mkdir -p gpurun_out/r02
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 5 2>&1 | tail -1
timeout 300 python tools/gpu_bringup.py tower --precision bf16 --n 3 2>&1 | tail -1
timeout 300 python tools/gpu_bringup.py profile --precision fp16 --n 128 2>&1 | grep -E "conv0|conv1|conv2" | head -3
timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 2>&1 | grep SPEED
timeout 900 python -m pytest tests -x -q -m gpu -k "parity or batch or ragged or u8 or rgb" 2>&1 | tail -3
