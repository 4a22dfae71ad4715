#!/bin/bash
# ncu --set full of ONE launch of the streaming stepper with ONE warp per scheduler (pop 592 = 148 SMs x 4, team 1):
# what a lone warp stalls on -- the latency of a trajectory, which bounds small populations and the tail of a shard.
# usage: tools/profile_single_warp.sh <output stem under gpurun_out/>
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TOSS_B200_TEAM=1
B="bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-microbench --pop 592"
python $B > gpurun_out/$1_plain.json     # must exit 0 without ncu first
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:stepper_kernel<\(bool\)1, \(int\)1, \(bool\)0' -s 3 -c 1 -o gpurun_out/$1 -f python $B > gpurun_out/$1_ncu.log 2>&1
tail -3 gpurun_out/$1_ncu.log
