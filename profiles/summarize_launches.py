"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi: continue
    name = re.sub(r'\(.*', '', r[ki]); v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] in ('ns', 'nsecond') else (v * 1000 if r[ui] in ('ms', 'msecond') else v)
    a = agg.setdefault(name, [0, 0.0, 1e9, 0]); a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
tot = sum(a[1] for a in agg.values())
print(f"{len(data)} launches, {tot:.1f} us total (serialised, cold-cache: compare shares, not absolutes)")
print(f"{'kernel':42s} {'n':>5s} {'total_us':>10s} {'avg_us':>8s} {'min_us':>8s} {'max_us':>8s} {'share':>6s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} {a[0]:5d} {a[1]:10.1f} {a[1]/a[0]:8.2f} {a[2]:8.2f} {a[3]:8.2f} {100*a[1]/tot:5.1f}%")
