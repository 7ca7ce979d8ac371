"""Group an ncu --csv launch list (gpu__time_duration.sum + launch__grid_size) by (kernel, grid): where a vocoder call's time goes."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
idc, kn, mn, mv = H.index("ID"), H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
rec = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv:
        continue
    d = rec.setdefault(r[idc], {"name": r[kn]})
    d[r[mn]] = float(r[mv].replace(",", ""))
agg = collections.OrderedDict()
tot = 0.0
for d in rec.values():
    us = d.get("gpu__time_duration.sum", 0.0) / 1e3
    key = (d["name"][d["name"].find("conv_umma"):][:60] if "conv_umma" in d["name"] else d["name"][:60], int(d.get("launch__grid_size", 0)), int(d.get("launch__shared_mem_per_block_dynamic", 0)) // 1024)
    a = agg.setdefault(key, [0.0, 0, 0.0])
    a[0] += us; a[1] += 1
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot += us
print(f"total {tot/1e3:.3f} ms over {len(rec)} launches")
for (name, grid, smem), (us, c, by) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{100*us/tot:6.2f}% {us/1e3:9.3f} ms  n={c:4d}  avg {us/c:8.1f} us  dram {by/max(us,1e-9)/1e3:7.0f} GB/s  grid {grid:7d} smem {smem:3d}K  {name}")
