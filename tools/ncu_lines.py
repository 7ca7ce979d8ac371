#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <object.o> <demangled-substring> <mangled-substring> [top_n]

ncu's CSV source page only carries SASS; this joins it (by instruction offset) with the line
table that `nvdisasm -g` prints for the same cubin (built with -lineinfo).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(obj, kernel_sub):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    table, cur, infn = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            infn = kernel_sub in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, obj, ksub, msub = sys.argv[1:5]
    topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    table = line_table(obj, msub)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data, take = None, [], False
    for r in rows:
        if r and r[0] == "Kernel Name":
            take = ksub in r[1] and hdr is None
            continue
        if r and r[0] == "Address":
            if take:
                hdr = r
            continue
        if take and hdr and len(r) == len(hdr):
            data.append(r)
        elif hdr and data and r and r[0] == "Kernel Name":
            break
    si = hdr.index("Warp Stall Sampling (All Samples)")
    base = min(int(r[0], 16) for r in data)
    by_line = defaultdict(int)
    by_line_ins = defaultdict(lambda: defaultdict(int))
    total = 0
    for r in data:
        off = int(r[0], 16) - base
        s = int(r[si] or 0)
        total += s
        loc, sass = table.get(off, (None, r[1]))
        by_line[loc] += s
        by_line_ins[loc][sass.split()[0] if sass else "?"] += s
    print(f"kernel *{ksub}*: {total} stall samples over {len(data)} instructions")
    src_cache = {}
    for loc, s in sorted(by_line.items(), key=lambda kv: -kv[1])[:topn]:
        text = ""
        if loc:
            f, n = loc
            for root in ("gsv-tts-lite_b200/csrc", "."):
                p = os.path.join(root, f)
                if os.path.exists(p):
                    src_cache.setdefault(p, open(p).read().splitlines())
                    if n - 1 < len(src_cache[p]):
                        text = src_cache[p][n - 1].strip()
                    break
        ins = ",".join(f"{k}:{v}" for k, v in sorted(by_line_ins[loc].items(), key=lambda kv: -kv[1])[:3])
        print(f"{100.0 * s / total:6.2f}%  {str(loc):34s} {text[:70]:70s} [{ins}]")


if __name__ == "__main__":
    main()
