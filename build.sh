#!/bin/bash
# Builds libstablemtl_sm100.so in-tree (sm_100a only). Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
SRC="stablemtl_b200/csrc/smtl_api.cu stablemtl_b200/csrc/smtl_gemm.cu stablemtl_b200/csrc/smtl_elem.cu stablemtl_b200/csrc/smtl_attn.cu"
nvcc --threads 4 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -o stablemtl_b200/libstablemtl_sm100.so $SRC "$@"
