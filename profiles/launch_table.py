"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0][:70] if "elementwise" not in r[ki] else r[ki][:230]
    v = float(r[vi].replace(',', '')); u = r[ui]
    us = v / 1000 if u.startswith('ns') else (v if u.startswith('us') else v * 1000)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot/1000:.2f} ms total (cold-cache, serialised: compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={n:4d}  {k}")
