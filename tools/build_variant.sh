#!/bin/bash
# A/B builds of the kernels: tools/build_variant.sh NAME [-DMACRO=..]... -> build/libslam2d_NAME.so
# (select it at run time with SLAM2D_B200_LIB=build/libslam2d_NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
cd slam-2d-lidar-scan_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -rdc=true -maxrregcount=128 -Xcompiler -fPIC -shared "$@" \
  -o ../../build/libslam2d_$name.so api.cu match.cu update.cu filter.cu
