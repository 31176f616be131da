cd /root/repo
mkdir -p gpurun_out/r02
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-latency > gpurun_out/r02/bench_${N}gpu_weak.log 2>&1; grep "^{" gpurun_out/r02/bench_${N}gpu_weak.log | cut -c1-200
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-latency > gpurun_out/r02/bench_1gpu_on_${N}box.log 2>&1; grep "^{" gpurun_out/r02/bench_1gpu_on_${N}box.log | cut -c1-200
