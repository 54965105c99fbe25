#!/usr/bin/env python
"""ncu launch list of ONE plain step (all kernels, whole batch per launch: DTMPC_CHUNKS=1) -> profiles/<name>.json:

    DTMPC_CHUNKS=1 ncu --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum,\
        dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'dtmpc|reduce_' -s <first> -c <n> \
        --csv --log-file gpurun_out/step.csv python tools/step_times.py --steps 1 ...
    python tools/roofline_counters.py gpurun_out/step.csv --batch 65536 --horizon 50 --out profiles/r02_step_counters.json

The JSON holds, per kernel family, the totals over the step and the per-unit constants bench.py's executed-work
roofline uses: fp64-pipe warp instructions, issued warp instructions and DRAM bytes per TILE-STEP (32 instances x one
horizon step; the line search per candidate), taken from the launches that ran with every instance active.
"""
import argparse, collections, csv, json, os, re

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--batch", type=int, required=True)
ap.add_argument("--horizon", type=int, required=True)
ap.add_argument("--out", required=True)
ap.add_argument("--note", default="")
ap.add_argument("--pipe-metric", default="sm__inst_executed_pipe_fp64.sum",
                help="arithmetic-pipe counter stored as fp64_inst (fp32 runs: sm__inst_executed_pipe_fma.sum)")
a = ap.parse_args()

rows = list(csv.reader(l for l in open(a.csv) if l.startswith('"')))
hdr, data = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
launches = collections.OrderedDict()
for r in data:
    key = int(r[ix["ID"]])
    d = launches.setdefault(key, {"name": r[ix["Kernel Name"]], "block": r[ix["Block Size"]], "grid": r[ix["Grid Size"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))


def family(name):
    for f in ("init", "ric", "fwd", "commit", "fwdc", "sens", "ift_accumulate", "reduce_partial", "reduce_final", "plant_step"):
        if re.search(rf"\b{f}_kernel", name):
            return f
    return "other"


tiles = (a.batch + 31) // 32
fam = collections.OrderedDict()
for d in launches.values():
    f = family(d["name"])
    e = fam.setdefault(f, {"launches": 0, "inst": 0.0, "fp64_inst": 0.0, "dram_bytes": 0.0, "time_us": 0.0, "first": None})
    e["launches"] += 1
    e["inst"] += d.get("smsp__inst_executed.sum", 0.0)
    e["fp64_inst"] += d.get(a.pipe_metric, 0.0)
    e["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    e["time_us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    if e["first"] is None:
        e["first"] = d
for e in fam.values():
    e.pop("first")
# per-unit constants from the launches that ran with every instance active: the first ric / fwdc launch after each
# init_kernel (first init = nominal OCP, second = auxiliary OCP, which also streams the tracked reference) and the
# sensitivity launches in order (auxiliary, then nominal)
per_unit = {}
ocp, n_sens = None, 0
for d in launches.values():
    f = family(d["name"])
    if f == "init":
        ocp = "n" if ocp is None else "a"
    key = None
    if f in ("init", "ric", "fwd", "commit", "fwdc") and ocp is not None:
        key = f"{f}_{ocp}"
    elif f == "sens":
        key = "sens_a" if n_sens == 0 else "sens_n"
        n_sens += 1
    if key is None or key in per_unit:
        continue
    units = tiles * a.horizon
    if f in ("fwd", "fwdc"):
        m = re.search(r"\((\d+), (\d+), 1\)", d["block"])
        units *= int(m.group(2))                      # one warp per candidate
    per_unit[key] = {"fp64_inst": d.get(a.pipe_metric, 0.0) / units,
                     "inst": d.get("smsp__inst_executed.sum", 0.0) / units,
                     "dram_bytes": (d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / units,
                     "time_us_first_launch": d.get("gpu__time_duration.sum", 0.0) / 1e3, "kernel": d["name"].split("(")[0]}
tot = {k: sum(e[k] for e in fam.values()) for k in ("inst", "fp64_inst", "dram_bytes", "time_us")}
out = {"source": os.path.basename(a.csv), "note": a.note, "pipe_metric": a.pipe_metric, "batch": a.batch, "horizon": a.horizon, "n_launches": len(launches),
       "step_totals": tot, "per_solve": {k: v / a.batch for k, v in tot.items()}, "families": fam,
       "per_tile_step": per_unit,
       "unit": "tile-step = 32 instances x 1 horizon step of one launch (line search: per candidate warp)"}
json.dump(out, open(a.out, "w"), indent=1)
print(json.dumps({"step_totals": tot, "per_solve": out["per_solve"], "per_tile_step": {k: {q: round(v, 1) if isinstance(v, float) else v for q, v in u.items()} for k, u in per_unit.items()}}, indent=1))
