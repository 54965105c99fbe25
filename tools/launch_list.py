"""Summarises an ncu launch list csv (gpu__time_duration.sum): per-kernel totals and the per-iteration timeline."""
import csv, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r['Metric Name'] == 'gpu__time_duration.sum']
seq = [(r['Kernel Name'].split('<')[0].replace('void dtmpc::', '').replace('void ', '')[:28], float(r['Metric Value'].replace(',', '')) / 1e3) for r in rows]
agg = collections.OrderedDict()
for n, t in seq:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {c:5d}x {100 * t / tot:5.1f}%  {n}")
print(f"total {tot:.1f} us over {len(seq)} launches")
if len(sys.argv) > 2:
    for n, t in seq: print(f"{n:28s} {t:9.1f}")
