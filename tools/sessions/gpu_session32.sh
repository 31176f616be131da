#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
ICP_TIMELINE=gpurun_out/tl_c35_1x1.bin timeout 200 python tools/gpu_bringup.py layer --name c35_1x1_288_256 --reps 10 --n 64 > gpurun_out/tl3.log 2>&1
timeout 300 python tools/gpu_bringup.py conv > gpurun_out/conv_cases.log 2>&1; echo "conv rc=$?"
grep -E "BAD|EXC" gpurun_out/conv_cases.log | head
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/tl3.log | grep LAYER
