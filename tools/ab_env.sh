#!/usr/bin/env bash
# step time of the in-tree library under different launch-structure settings (chunks per call, steps in flight)
cd "$(dirname "$0")/.."
for cfg in "4 2" "2 2" "1 2" "8 2" "4 3" "2 3" "4 2"; do
  set -- $cfg
  DTMPC_CHUNKS=$1 DTMPC_BENCH_BUFFERS=$2 python bench.py --steps 10 --warmup 3 --preheat-seconds 1 --no-cpu-baseline --no-extras ${BENCH_ARGS:-} > gpurun_out/env_$1_$2.json 2> gpurun_out/env_$1_$2.err
  python -c "
import json;d=json.loads(open('gpurun_out/env_$1_$2.json').read().strip().splitlines()[-1]);print('[env] chunks $1 buffers $2:', round(d['ms_per_step'],3), 'seq', round(d['sequential']['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done
