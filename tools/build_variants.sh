#!/bin/bash
# developer experiment: build stepper variants into gpurun_out-independent .so files under build/variants
set -e
cd "$(dirname "$0")/../toss_b200/csrc"
mkdir -p ../../build/variants
build() { # name flags...
  name=$1; shift
  nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" -shared -o ../../build/variants/libtoss_$name.so api.cu gravity_kernels.cu stepper.cu fitness.cu -lcudart 2> ../../build/variants/$name.log
  grep -A2 "stepper_kernelILb1" ../../build/variants/$name.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name"
}
build base
build u1 -DK2_UNROLL=1
build w16u1 -DK2_WARPS=16 -DK2_MINBLOCKS=1 -DK2_STAGES=4 -DK2_UNROLL=1
build w16 -DK2_WARPS=16 -DK2_MINBLOCKS=1 -DK2_STAGES=4
build st2 -DK2_STAGES=2
