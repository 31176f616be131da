cd /root/repo
mkdir -p gpurun_out/r02
run() { local name=$1; shift; local t=$1; shift; echo "=== $name"; timeout "$t" "$@" > gpurun_out/r02/$name.log 2>&1; echo "=== $name rc=$?"; }
run pytest_gpu 900 python -m pytest tests -q -m gpu
tail -3 gpurun_out/r02/pytest_gpu.log
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
tail -1 gpurun_out/r02/smoke.log
run bench_ref 400 python bench.py --impl reference --steps 8 --warmup 3
run bench_fp16 900 python bench.py --gpus 1 --steps 30 --warmup 5
grep "^{" gpurun_out/r02/bench_fp16.log | cut -c1-300
run bench_bf16 900 python bench.py --gpus 1 --steps 30 --warmup 5 --precision bf16 --no-cpu-baseline
grep "^{" gpurun_out/r02/bench_bf16.log | cut -c1-200
run http 900 python tools/http_bench.py --workers 8 --procs 8 --threads 4 --requests 6000
grep "^{" gpurun_out/r02/http.log | cut -c1-300
run profile_b128 300 python tools/gpu_bringup.py profile --n 128 --precision fp16
run profile_b1 300 python tools/gpu_bringup.py profile --n 1 --precision fp16 --reps 20
run speed_b1 300 python tools/gpu_bringup.py speed --precision fp16 --n 1 --graph --iters 300 --u8
tail -1 gpurun_out/r02/speed_b1.log
run ncu_full_all 2400 ncu --set full --clock-control none --import-source on -s 240 -c 80 -o gpurun_out/r02/all_full python tools/gpu_bringup.py speed --n 128 --iters 1 --precision fp16 --u8
ncu -i gpurun_out/r02/all_full.ncu-rep --page raw --csv > gpurun_out/r02/all_full_raw.csv 2>/dev/null
ls -la gpurun_out/r02/
