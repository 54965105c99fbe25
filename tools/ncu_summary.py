#!/usr/bin/env python
"""Text summary of an `ncu --set full` report: launch shape, occupancy, pipe / issue utilisation, cache hit rates, DRAM
bytes, stall reasons per issue, and -- if the report carries source counters (--import-source on) -- the dynamic opcode
mix per loop (grouped by execution count) and the hottest stall locations.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--kernel-index 0] [--units NAME=COUNT ...]
"""
import argparse, collections, csv, io, re, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--units", nargs="*", default=[], help="name=min_exec:max_exec:unit_count groups for the opcode mix")
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("----")
    print("Kernel Name".ljust(72), d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            unit = dict(zip(hdr, rows[1])).get(k, "")
            print(k.ljust(72), d[k], unit)
    print("stall reasons (cycles per issued instruction):")
    for k, v in sorted(d.items()):
        if "average_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                if float(v) > 0.1:
                    print("  " + k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "").ljust(28), f"{float(v):.3f}")
            except ValueError:
                pass
src = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 3 and "Instructions Executed" in srows[1]:
    h = srows[1]; data = srows[2:]; ix = {k: i for i, k in enumerate(h)}
    tot = sum(float(r[ix["# Samples"]] or 0) for r in data if len(r) > ix["# Samples"])
    for g in a.units:
        name, spec = g.split("=")
        lo, hi, n = (float(x) for x in spec.split(":"))
        c = collections.Counter()
        for r in data:
            ex = float(r[ix["Instructions Executed"]] or 0)
            if lo <= ex < hi:
                op = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip()).split()[0].split(".")[0]
                c[op] += ex / n
        fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
        print(f"dynamic opcode mix, {name} ({n:.0f} units): {sum(c.values()):.1f} issued per unit, fp64 pipe {fp64:.1f}")
        print("    " + "  ".join(f"{k}:{v:.1f}" for k, v in c.most_common(24)))
    print("hottest stall locations (share of all warp-stall samples; the instruction shown WAITS for its inputs):")
    for r in sorted(data, key=lambda r: -float(r[ix["# Samples"]] or 0))[:12]:
        s = float(r[ix["# Samples"]])
        st = {k: float(r[ix[k]] or 0) for k in ("stall_long_sb", "stall_wait", "stall_math", "stall_short_sb", "stall_not_selected", "stall_barrier")}
        print(f"  {s / tot * 100:5.2f}%  {max(st, key=st.get)[6:]:>12}  executed {float(r[ix['Instructions Executed']]):>9.0f}  {r[ix['Source']].strip()[:64]}")
