#!/bin/bash
# tuning helper: build libort with extra -D flags into variants_tmp_<name>.so (see tools/variants.sh)
# usage: tools/build_variant.sh <name> [-DORT_MIN_BLOCKS=4 ...]
set -e
name=$1; shift
cd "$(dirname "$0")/../opticalraytrace_b200/csrc"
NV=/usr/local/cuda/bin/nvcc
$NV -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -Xptxas -v "$@" \
    -c -o /tmp/ort_cuda_$name.o ort_cuda.cu 2> /tmp/ptxas_$name.log
grep -A2 "ort_trace_kernelILi2ELi1ELi0EdE" /tmp/ptxas_$name.log | grep -E "spill|registers" | head -2
$NV -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ../../variants_tmp_$name.so ort_host.o /tmp/ort_cuda_$name.o ort_bpm.o -ldl
