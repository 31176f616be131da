#!/bin/bash
# Round-end artefacts: tests, smoke, bench (both arms), ncu launch list, ncu DRAM traffic of every conv launch.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session16.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/session16.log
  tail -n 40 "gpurun_out/$name.log" >> gpurun_out/session16.log
}
run pytest_gpu 900 python -m pytest tests -x -q -m gpu
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench_ref 400 python bench.py --impl reference --steps 8 --warmup 3
run bench1 600 python bench.py --gpus 1 --steps 30 --warmup 5
run bench1_fp16 600 python bench.py --gpus 1 --steps 30 --warmup 5 --precision fp16 --no-cpu-baseline
run profile_b128 300 python tools/gpu_bringup.py profile --n 128 --precision bf16
run profile_b1 300 python tools/gpu_bringup.py profile --n 1 --precision bf16 --reps 20
run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python tools/gpu_bringup.py speed --n 128 --iters 1
run ncu_traffic 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_ -s 84 -c 84 --csv --log-file gpurun_out/conv_traffic_r01_final.csv python tools/gpu_bringup.py speed --n 128 --iters 1
run ncu_full_top 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_persistent -s 36 -c 6 -o gpurun_out/persist_final python tools/gpu_bringup.py speed --n 128 --iters 1
grep -E "^===|passed|failed|smoke|PROFILE|^\{" gpurun_out/session16.log | cut -c1-2500
