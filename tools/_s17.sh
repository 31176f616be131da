cd /root/repo
mkdir -p gpurun_out/r02
timeout 900 python tools/http_bench.py --workers 8 --procs 8 --threads 4 --requests 6000 > gpurun_out/r02/http.log 2>&1; grep "^{" gpurun_out/r02/http.log | cut -c1-900
timeout 900 python tools/http_bench.py --workers 8 --procs 8 --threads 8 --requests 6000 --max-batch 64 --window-ms 3 > gpurun_out/r02/http64.log 2>&1; grep "^{" gpurun_out/r02/http64.log | cut -c1-600
timeout 900 python tools/http_bench.py --workers 8 --procs 8 --threads 4 --requests 3000 --decode owner > gpurun_out/r02/http_owner.log 2>&1; grep "^{" gpurun_out/r02/http_owner.log | cut -c1-400
timeout 600 python -m pytest tests -q -m gpu -k serving 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_persistent -s 130 -c 3 -o gpurun_out/r02/persist_full python tools/gpu_bringup.py speed --n 128 --iters 1 --precision fp16 > gpurun_out/r02/ncu_full.log 2>&1; tail -3 gpurun_out/r02/ncu_full.log
ls -la gpurun_out/r02/*.ncu-rep
