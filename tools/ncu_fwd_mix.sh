#!/usr/bin/env bash
# usage (on the GPU box, via gpurun): tools/ncu_fwd_mix.sh <tag> [kernel-regex]
set -e
TAG=$1; KR=${2:-fwdc_kernel}
CMD="env DTMPC_CHUNKS=1 python bench.py --batch 65536 --steps 1 --warmup 1 --no-cpu-baseline --mode ${MODE:-full} --sequential"
$CMD > gpurun_out/plain_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$KR" -s 3 -c 1 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
