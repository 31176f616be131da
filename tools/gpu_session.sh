#!/bin/bash
# One parameterised GPU session (replaces the numbered one-off scripts of round 1):
#   gpurun --timeout 1500 -- bash tools/gpu_session.sh <step> [<step> ...]
# Every step writes gpurun_out/<step>.log, bounded by its own `timeout`; a summary is printed at the end.
# Steps: tests | tests_fast | smoke | bench | bench_bf16 | bench_ref | profile | profile_b1 | speed | speed_b1 |
#        ncu_launches | ncu_traffic | ncu_kernels | ncu_full | sass | http | env:<VAR=VALUE> (applies to later steps) |
#        py:<script args>  (python tools/gpu_bringup.py <args>)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export ICP_MBAR_TIMEOUT_MS=${ICP_MBAR_TIMEOUT_MS:-3000}     # bring-up: a protocol bug traps after ~3 s instead of ~60 s
SUM=gpurun_out/session.log
: > $SUM
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name: $*" >> $SUM
  local t0=$(date +%s)
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "=== $name rc=$rc $(( $(date +%s) - t0 ))s" >> $SUM
  grep -E "passed|failed|error|Error|smoke|PROFILE|SPEED|LAYER|^\{" "gpurun_out/$name.log" | tail -n 12 | cut -c1-1800 >> $SUM
}
NCU_METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
i=0
for step in "$@"; do
  i=$((i+1))
  case "$step" in
    env:*) export "${step#env:}"; echo "=== env ${step#env:}" >> $SUM ;;
    tests) run tests 1500 python -m pytest tests -x -q -m gpu ;;
    tests_fast) run tests_fast 900 python -m pytest tests -x -q -m gpu -k "not every_tower_layer_shape and not edge_cases and not opt_in" ;;
    smoke) run smoke 300 python -c "import __graft_entry__ as g; g.smoke()" ;;
    bench) run bench_$i 900 python bench.py --gpus 1 --steps 30 --warmup 5 ;;
    bench_quick) run bench_quick_$i 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-latency --sustain-seconds 0 ;;
    bench_bf16) run bench_bf16 900 python bench.py --gpus 1 --steps 30 --warmup 5 --precision bf16 --no-cpu-baseline ;;
    bench_ref) run bench_ref 400 python bench.py --impl reference --steps 8 --warmup 3 ;;
    profile) run profile_$i 300 python tools/gpu_bringup.py profile --n 128 --precision fp16 ;;
    profile_b1) run profile_b1_$i 300 python tools/gpu_bringup.py profile --n 1 --precision fp16 --reps 20 ;;
    speed) run speed_$i 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph ;;
    speed_b1) run speed_b1_$i 300 python tools/gpu_bringup.py speed --precision fp16 --n 1 --graph --iters 200 ;;
    ncu_launches) run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python tools/gpu_bringup.py speed --n 128 --iters 1 --precision fp16 ;;
    ncu_kernels) run ncu_kernels 1200 ncu --metrics $NCU_METRICS --clock-control none -c 800 --csv --log-file gpurun_out/kernels.csv python tools/gpu_bringup.py speed --n 128 --iters 1 --precision fp16 ;;
    ncu_full) run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-conv_igemm_persistent} -s ${NCU_SKIP:-230} -c ${NCU_COUNT:-3} -o gpurun_out/${NCU_OUT:-full} python tools/gpu_bringup.py speed --n 128 --iters 1 --precision fp16 ;;
    http) run http 900 python tools/http_bench.py --workers 8 --clients 32 --requests 2000 ;;
    py:*) run py_$i 900 python tools/gpu_bringup.py ${step#py:} ;;
    sh:*) run sh_$i 900 bash -c "${step#sh:}" ;;
    *) echo "unknown step $step" >> $SUM ;;
  esac
done
cat $SUM
