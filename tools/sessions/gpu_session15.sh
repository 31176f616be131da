#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session15.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/session15.log
  tail -n 60 "gpurun_out/$name.log" >> gpurun_out/session15.log
}
run conv_cases 300 python tools/gpu_bringup.py conv
run tower_fp16 300 python tools/gpu_bringup.py tower --precision fp16 --n 2
run profile_b128 300 python tools/gpu_bringup.py profile --n 128 --precision bf16
run nopair_profile 300 env ICP_PAIR=0 python tools/gpu_bringup.py profile --n 128 --precision bf16
run speed_b128_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 128 --graph
run speed_b1_graph 300 python tools/gpu_bringup.py speed --precision bf16 --n 1 --graph --iters 50
run pytest_gpu 900 python -m pytest tests -x -q -m gpu
run bench1 600 python bench.py --gpus 1 --steps 20 --warmup 3
grep -E "^===|CASE.*(BAD|EXC)|max \|dprob|PROFILE|SPEED|passed|failed|icp:|^\{" gpurun_out/session15.log | cut -c1-900 | tail -40
paste <(grep "^OP " gpurun_out/profile_b128.log | awk '{for(i=3;i<=NF;i++){ if($i ~ /^ms=/){ms=$(i+1)} } printf "%-58s %s\n", substr($2,1,58), ms}') <(grep "^OP " gpurun_out/nopair_profile.log | awk '{for(i=3;i<=NF;i++){ if($i ~ /^ms=/){ms=$(i+1)} } print ms}')
