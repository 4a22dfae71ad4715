#!/bin/bash
# developer experiment: build stepper variants into .so files under build/variants (tools/variant_bench.py times them)
set -e
cd "$(dirname "$0")/../toss_b200/csrc"
mkdir -p ../../build/variants
build() { # name flags...
  name=$1; shift
  nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" -shared -o ../../build/variants/libtoss_$name.so api.cu pairs.cu gravity_kernels.cu stepper.cu fitness.cu sharding.cu gaco.cu -lcudart 2> ../../build/variants/$name.log
  grep -A2 "stepper_kernelILb1ELi1ELb0ELb0" ../../build/variants/$name.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name"
}
for v in "$@"; do
  case $v in
    base)   build base ;;
    unroll2) build unroll2 -DK2_UNROLL=2 ;;
    st4)    build st4 -DK2_STAGES=4 ;;
    t64s6)  build t64s6 -DPAIR_TILE=64 -DK2_STAGES=6 ;;
    t64s8)  build t64s8 -DPAIR_TILE=64 -DK2_STAGES=8 ;;
    w12)    build w12 -DK2_WARPS=12 ;;
    w14)    build w14 -DK2_WARPS=14 ;;
    w10)    build w10 -DK2_WARPS=10 ;;
    w8)     build w8 -DK2_WARPS=8 ;;
    w12u2)  build w12u2 -DK2_WARPS=12 -DK2_UNROLL=2 ;;
    w20)    build w20 -DK2_WARPS=20 ;;
    w24)    build w24 -DK2_WARPS=24 ;;
    hint)   build hint -DK2_WAIT_HINT=20000 ;;
    nocompact) build nocompact -DK2_NO_COMPACT ;;
    w8x2s2) build w8x2s2 -DK2_WARPS=8 -DK2_MINBLOCKS=2 -DK2_STAGES=2 -DK2_NO_TEAMBUF ;;
    s4nt)   build s4nt -DK2_STAGES=4 -DK2_NO_TEAMBUF ;;
    k1mb5)  build k1mb5 -DK1_MINBLOCKS=5 ;;
    k1mb6)  build k1mb6 -DK1_MINBLOCKS=6 ;;
    k1mb3u2) build k1mb3u2 -DK1_MINBLOCKS=3 -DK1_UNROLL=2 ;;
    k1mb4u2) build k1mb4u2 -DK1_MINBLOCKS=4 -DK1_UNROLL=2 ;;
    prof)   build prof -DK2_PROFILE ;;
    prof4)  build prof4 -DK2_PROFILE -DK2_STAGES=4 ;;
    *) echo "unknown variant $v"; exit 1 ;;
  esac
done
