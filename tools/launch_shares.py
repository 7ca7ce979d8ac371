"""Summarise an ncu --csv launch list (gpu__time_duration.sum): share of device time per kernel name."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
kn, mv, mu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
tot = 0.0
n = 0
for r in rows[hdr + 1:]:
    if len(r) <= mv:
        continue
    n += 1
    if n <= skip:
        continue
    v = float(r[mv].replace(",", ""))
    u = r[mu]
    us = v / 1e3 if u == "ns" else v * (1e3 if u == "ms" else 1.0) if u in ("ms", "us", "usecond") else v / 1e3
    name = r[kn][:90]
    a = agg.setdefault(name, [0.0, 0])
    a[0] += us; a[1] += 1
    tot += us
print(f"total {tot/1e3:.3f} ms over {sum(a[1] for a in agg.values())} launches")
for name, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100*us/tot:6.2f}% {us/1e3:9.3f} ms  n={c:5d}  avg {us/c:8.1f} us  {name}")
