#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_bringup.py tower --precision fp16 --n 2 > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|dprob|out of bounds|misaligned" gpurun_out/sanitizer_memcheck.log | head -20
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_bringup.py conv --only halo > gpurun_out/sanitizer_conv.log 2>&1; echo "memcheck conv rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/sanitizer_conv.log | head -10
