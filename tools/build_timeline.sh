#!/bin/bash
# Builds tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so with the clock64 timeline stamps compiled in
# (use: ICP_AB_LIB=<that file> ICP_TIMELINE=<out.bin> python tools/gpu_bringup.py layer --name ... ; tools/timeline.py)
cd "$(dirname "$0")/.."
mkdir -p tensorflow_web_deploy_b200/_ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
  -DICP_TIMELINE_BUILD -Iinclude -o tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so \
  tensorflow_web_deploy_b200/csrc/engine.cu -lcuda -lnvjpeg_static -lculibos
