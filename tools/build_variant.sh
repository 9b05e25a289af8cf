#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG=..]... : builds variants/libhfb200_NAME.so (kernel experiments; travels to the GPU box, git-ignored)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" -o variants/libhfb200_$name.so hyperfridge-r0_b200/csrc/hfb200.cu
echo variants/libhfb200_$name.so
