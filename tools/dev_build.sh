#!/bin/bash
# Experiment builds of liblyon_cuda (one tile-kernel instantiation; optional -D flags), written next to the product .so:
#   tools/dev_build.sh phase -DLYON_PHASE_TIMING     ->  lyon_b200/liblyon_cuda_phase.so
# Use with LYON_CUDA_LIB=lyon_b200/liblyon_cuda_<tag>.so.
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -shared \
     -DLYON_DEV_BUILD "$@" -o lyon_b200/liblyon_cuda_${tag}.so lyon_b200/csrc/lyon_cuda.cu
