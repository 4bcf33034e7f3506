"""Summarise an ncu --csv launch list: per kernel (last construct only) time, instructions, DRAM bytes, issue utilisation."""
import csv, collections, sys
path = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.75
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    try: v = float(row['Metric Value'].replace(',', ''))
    except Exception: continue
    agg.setdefault((int(row['ID']), row['Kernel Name'][:78]), {})[row['Metric Name']] = v
ids = sorted(i for i, _ in agg)
cut = ids[0] + int((ids[-1] - ids[0] + 1) * frac)
per = collections.OrderedDict()
for (i, k), d in agg.items():
    if i < cut: continue
    p = per.setdefault(k, collections.Counter()); p['n'] += 1
    for m, v in d.items(): p[m] += v
print("%-80s %3s %9s %9s %8s %8s %6s" % ("kernel", "n", "time us", "inst M", "rd MB", "wr MB", "iss%"))
tot = 0
for k, p in per.items():
    n = p['n']; tot += p['gpu__time_duration.sum']
    print("%-80s %3d %9.1f %9.1f %8.0f %8.0f %6.1f" % (k, n, p['gpu__time_duration.sum'] / 1e3, p['smsp__inst_executed.sum'] / 1e6, p['dram__bytes_read.sum'] / 1e6,
                                                     p['dram__bytes_write.sum'] / 1e6, p['smsp__issue_active.avg.pct_of_peak_sustained_active'] / n))
print("total kernel time %.1f us" % (tot / 1e3))
