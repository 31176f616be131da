cd /root/repo
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 1 2>&1 | tail -1
timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 3 2>&1 | tail -1
for N in 1 2 4 8 16; do
for NT in 0 16 32 64; do
ICP_NARROW_NTILE=$NT timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n $N --graph --iters 30 --u8 2>&1 | grep SPEED | sed "s/^/narrow=$NT /" | cut -c1-100
done
done
timeout 900 python -m pytest tests -x -q -m gpu -k "parity or batch or ragged or u8 or rgb or thread" 2>&1 | tail -3
