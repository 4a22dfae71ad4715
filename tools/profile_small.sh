#!/bin/bash
# ncu --set full of ONE launch of the RESIDENT stepper on a small population (BASELINE configs[0] / configs[2] shapes):
# where the latency of one RHS evaluation goes.  usage: tools/profile_small.sh <stem> <mesh> <pop>
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-microbench --mesh $2 --pop $3"
python $B > gpurun_out/$1_plain.json     # must exit 0 without ncu first
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:stepper_kernel<\(bool\)0' -s 3 -c 1 -o gpurun_out/$1 -f python $B > gpurun_out/$1_ncu.log 2>&1
tail -2 gpurun_out/$1_ncu.log
