#!/bin/bash
# Experiment builds of the library for tools/ab.py: bash tools/build_variant.sh <name> [-DFLAG=VALUE ...]  ->  build_exp/<name>.so
set -e
NAME=$1; shift
mkdir -p build_exp
nvcc -shared -std=c++17 -O3 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -o build_exp/$NAME.so unipeak_b200/csrc/upb_api.cu unipeak_b200/csrc/upb_multi.cu
echo build_exp/$NAME.so
