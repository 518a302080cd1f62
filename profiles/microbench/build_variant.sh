#!/bin/bash
# build_variant.sh NAME [-DFLAG=VALUE ...] : compile the library with extra defines into profiles/microbench/variants/NAME.so
# (A/B runs: AX3D_LIB=profiles/microbench/variants/NAME.so python bench.py ...)
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
nvcc -shared -Xcompiler -fPIC -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ftz=true -Xptxas -O3 \
  -DAX3D_WITH_NCCL "$@" -o profiles/microbench/variants/$name.so axisem3d_b200/csrc/api.cu -lnccl
