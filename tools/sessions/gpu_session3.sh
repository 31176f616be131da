#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/session3.log
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "=== $name rc=$?" | tee -a gpurun_out/session3.log
  tail -n 40 "gpurun_out/$name.log" >> gpurun_out/session3.log
}
run pytest_gpu 1200 python -m pytest tests -x -q -m gpu
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench1 600 python bench.py --gpus 1 --steps 20 --warmup 3
run bench_ref 600 python bench.py --impl reference --steps 3 --warmup 3
tail -n 120 gpurun_out/session3.log
