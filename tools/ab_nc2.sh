#!/usr/bin/env bash
# experiment: 2 candidates per thread (DTMPC_FWD_NC=2) at several register caps vs 1 candidate per thread (separate kernels)
cd "$(dirname "$0")/.."
M=smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum
for v in sep_nc1:1 sep_nc2_r96:2 sep_nc2_r112:2 sep_nc2_r128:2; do
  lib=${v%%:*}; nc=${v##*:}
  DTMPC_FWD_NC=$nc DTMPC_LIB=$PWD/variants/$lib.so python bench.py --steps 10 --warmup 3 --preheat-seconds 1 --no-cpu-baseline --no-extras > gpurun_out/nc_$lib.json 2> gpurun_out/nc_$lib.err
  python -c "
import json;d=json.loads(open('gpurun_out/nc_$lib.json').read().strip().splitlines()[-1]);print('[nc] $lib', round(d['ms_per_step'],3), round(d['sequential']['ms_per_step'],3))"
  DTMPC_FWD_NC=$nc DTMPC_LIB=$PWD/variants/$lib.so DTMPC_CHUNKS=1 ncu --metrics $M --clock-control none -k regex:'fwd_kernel' -c 3 --csv --log-file gpurun_out/nc_${lib}_inst.csv python tools/step_times.py --steps 1 > /dev/null 2>&1
  grep -h "fwd_kernel" gpurun_out/nc_${lib}_inst.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | head -9
done
