"""Reads gpurun_out/prof_<tag>.ncu-rep (here, no GPU): key metrics, stall reasons and the dynamic opcode mix."""
import csv, re, collections, subprocess, sys, io
tag = sys.argv[1]; per = float(sys.argv[2]) if len(sys.argv) > 2 else 65536 * 6 * 50 / 32
rep = f"gpurun_out/prof_{tag}.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum']
for k in keys:
    if k in hdr:
        i = hdr.index(k); print(f"{k:70s} {vals[i][:44]:>24s} {units[i]}")
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled') and float(vals[i] or 0) > 0.08:
        print(f"  stall {h[34:-23]:30s} {float(vals[i]):8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ia, ii, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    try: n = int(r[ii])
    except Exception: continue
    m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[ia].strip()); op = m.group(2) if m else '?'
    tot += n; ops[op] += n; samp[op] += int(r[isamp] or 0)
print(f"total warp-instr {tot}  per warp-unit {tot / per:.1f}")
print("  " + "  ".join(f"{op}:{n / per:.1f}" for op, n in ops.most_common(22)))
