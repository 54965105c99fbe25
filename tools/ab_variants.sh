#!/usr/bin/env bash
# A/B of library builds on the GPU box: for each variants/<name>.so, the bench step time (graph replay, two steps in
# flight and sequential) and, with NCU=1, executed-instruction counts of the solver kernels (whole batch in one launch).
#   tools/ab_variants.sh name1 name2 ...      -> gpurun_out/ab_<name>.json, gpurun_out/ab_<name>_inst.csv
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  lib="$PWD/variants/$v.so"
  [ -f "$lib" ] || { echo "missing $lib"; continue; }
  DTMPC_LIB="$lib" python bench.py --steps 10 --warmup 3 --warmup-seconds 1 --no-cpu-baseline ${BENCH_ARGS:-} \
      > "gpurun_out/ab_$v.json" 2> "gpurun_out/ab_$v.err"
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{v}.json").read().strip().splitlines()[-1])
    print(f"[ab] {v:12s} ms/step {d['ms_per_step']:.3f}  sequential {d['sequential']['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  "
          f"kernel_ms {d['roofline']['kernel_ms']}  iters {d['roofline']['iters_per_solve']:.3f} bad {d['nonfinite_instances']}")
except Exception as e:
    print(f"[ab] {v}: failed ({e})"); print(open(f"gpurun_out/ab_{v}.err").read()[-2000:])
PY
  if [ "${NCU:-0}" = "1" ]; then
    DTMPC_LIB="$lib" DTMPC_CHUNKS=1 ncu --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -k regex:'fwdc?_kernel|ric_kernel|commit_kernel|sens_kernel|init_kernel' -c ${NCU_COUNT:-16} --csv \
      --log-file "gpurun_out/ab_${v}_inst.csv" python tools/step_times.py --steps 1 ${STEP_ARGS:-} > "gpurun_out/ab_${v}_ncu.log" 2>&1
  fi
done
