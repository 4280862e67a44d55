#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG ...]  -> build/libspf_NAME.so (experiment builds, git-ignored)
name=$1; shift
mkdir -p build
nvcc -std=c++17 -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -shared -o build/libspf_$name.so spf_b200/csrc/capi.cu spf_b200/csrc/muxgen.o 2> build/ptxas_$name.log || { cat build/ptxas_$name.log; exit 1; }
grep -A2 pbs_kernel build/ptxas_$name.log | tail -2
