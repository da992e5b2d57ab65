#!/bin/bash
# Development: A/B library of assign_f16.cu with extra nvcc flags -> gpurun_variants/libvqb200_<name>.so
# usage: tools/build_variant.sh <name> [nvcc flags...]; run with VQB200_LIB_PATH=$PWD/gpurun_variants/libvqb200_<name>.so
set -e
name=$1; shift
pkg=$(ls -d "$(dirname "$0")"/../bridging*_b200)
mkdir -p "$(dirname "$0")/../gpurun_variants"
out="$(dirname "$0")/../gpurun_variants/libvqb200_$name.so"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 "$@" -c "$pkg/csrc/assign_f16.cu" -o "/tmp/assign_f16_$name.o"
objs=$(ls "$pkg"/build/*.o | grep -v assign_f16.o)
nvcc -shared -o "$out" $objs "/tmp/assign_f16_$name.o" -lcudart
echo "$out"
