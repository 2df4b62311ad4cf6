#!/bin/bash
# Builds libstablemtl_sm100.so in-tree (sm_100a only).  Same recipe as __graft_entry__.build(): every translation unit is
# compiled in parallel (smtl_gemm.cu twice: host code + fp16 kernels, and -DSMTL_GEMM_BF16_PART the bf16 kernels), then linked.
set -e
cd "$(dirname "$0")"
OBJ=$(mktemp -d)
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
pids=()
for u in smtl_api smtl_gemm smtl_elem smtl_attn; do
  nvcc $FLAGS -c stablemtl_b200/csrc/$u.cu -o $OBJ/$u.o "$@" & pids+=($!)
done
nvcc $FLAGS -DSMTL_GEMM_BF16_PART -diag-suppress=177 -c stablemtl_b200/csrc/smtl_gemm.cu -o $OBJ/smtl_gemm_bf16.o "$@" & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o stablemtl_b200/libstablemtl_sm100.so $OBJ/*.o
rm -rf $OBJ
