"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hdr_i]; data = rows[hdr_i + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for r in data:
    if len(r) <= vi: continue
    name = re.sub(r'\(.*', '', r[ki]); v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg[name][0] += 1; agg[name][1] += v; tot += v
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={n:4d} avg={t/n:8.1f}  {k[:100]}")
