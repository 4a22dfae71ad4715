#!/bin/bash
# ncu --set full of ONE launch of the streaming stepper (not its mascon-proxy instantiation), pop 16384 on lp.
# usage: tools/profile_stepper.sh <output stem under gpurun_out/>
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-microbench --pop 16384"
python $B > gpurun_out/$1_plain.json     # must exit 0 without ncu first
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:stepper_kernel<\(bool\)1, \(int\)1, \(bool\)0' -s 3 -c 1 -o gpurun_out/$1 -f python $B > gpurun_out/$1_ncu.log 2>&1
tail -3 gpurun_out/$1_ncu.log
