cd /root/repo
mkdir -p gpurun_out/r02
run() { local name=$1; shift; local t=$1; shift; echo "=== $name"; timeout "$t" "$@" > gpurun_out/r02/$name.log 2>&1; echo "=== $name rc=$?"; }
run bench_ref 400 python bench.py --impl reference --steps 8 --warmup 3
run bench_fp16 900 python bench.py --gpus 1 --steps 30 --warmup 5
grep "^{" gpurun_out/r02/bench_fp16.log | cut -c1-300
ICP_AB_LIB=tensorflow_web_deploy_b200/_ab/lib_base.so run bench_fp16_base 900 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
grep "^{" gpurun_out/r02/bench_fp16_base.log | cut -c1-300
run bench_fp16_again 900 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
grep "^{" gpurun_out/r02/bench_fp16_again.log | cut -c1-300
run bench_bf16 900 python bench.py --gpus 1 --steps 30 --warmup 5 --precision bf16 --no-cpu-baseline
grep "^{" gpurun_out/r02/bench_bf16.log | cut -c1-200
run http 900 python tools/http_bench.py --workers 8 --procs 8 --threads 4 --requests 6000
run profile_b128 300 python tools/gpu_bringup.py profile --n 128 --precision fp16
run profile_b1 300 python tools/gpu_bringup.py profile --n 1 --precision fp16 --reps 20
run speed_b1 300 python tools/gpu_bringup.py speed --precision fp16 --n 1 --graph --iters 300 --u8
run ncu_full_all 2400 ncu --set full --clock-control none -s 240 -c 80 -o /tmp/all_full python tools/gpu_bringup.py speed --n 128 --iters 1 --precision fp16 --u8
ncu -i /tmp/all_full.ncu-rep --page raw --csv > gpurun_out/r02/all_full_raw.csv 2>/dev/null
du -sh gpurun_out
