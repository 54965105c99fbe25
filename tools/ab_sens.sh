#!/usr/bin/env bash
cd "$(dirname "$0")/.."
for v in "$@"; do
  DTMPC_LIB=$PWD/variants/$v.so DTMPC_CHUNKS=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sens_kernel' -c 2 --csv --log-file gpurun_out/sens_$v.csv python tools/step_times.py --steps 1 > /dev/null 2>&1
  echo "[sens] $v $(grep gpu__time gpurun_out/sens_$v.csv | awk -F'","' '{printf "%s us  ", $NF}' | tr -d '"')"
done
