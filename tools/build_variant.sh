#!/bin/bash
# usage: tools/build_variant.sh NAME [extra nvcc -D flags...]   -> build/variants/libNAME.so
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -Xptxas -v "$@" \
  -o build/variants/lib$name.so bls_bn254_b200/csrc/capi.cu 2> build/variants/$name.ptxas.log
grep -A2 -E "k_miller2_fixed|k_final_exp|k_hash_g1" build/variants/$name.ptxas.log | grep -E "registers|stack" | paste - - | sed "s/^/$name: /" | cut -c1-230
