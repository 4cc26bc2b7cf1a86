#!/usr/bin/env python
"""Attribute an ncu source-page CSV (SASS rows) to source lines using nvdisasm -g line markers of the same cubin.

  ncu -i rep.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all lib.so; nvdisasm -g -c lib.sm_100a.cubin > dis.txt
  tools/ncu_lines.py src.csv dis.txt kernel_substring [top]
"""
import collections
import csv
import re
import sys

src_csv, dis_txt, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: offset -> (file, line) for the kernel's section (innermost inlined location)
line_of = {}
cur = None
infn = False
for ln in open(dis_txt):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln) or re.match(r'\s*\.text\.(\S+):', ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, iinst, ithr, isamp = (hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"),
                          hdr.index("# Samples"))
data = [r for r in rows[2:] if len(r) > isamp]


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def addr(x):
    return int(x, 16) if x.lower().startswith('0x') else int(x)


base = addr(data[0][ia])
agg = collections.defaultdict(lambda: [0, 0, 0])
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
stall_tot = collections.Counter()
for r in data:
    key = line_of.get(addr(r[ia]) - base, ('?', 0))
    g = agg[key]
    g[0] += num(r[iinst])
    g[1] += num(r[ithr])
    g[2] += num(r[isamp])
    for i in stall_cols:
        stall_tot[hdr[i]] += num(r[i])
ti = sum(g[0] for g in agg.values())
ts = sum(g[2] for g in agg.values())
print(f"total warp-inst {ti} samples {ts}")
byfile = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), g in agg.items():
    for k in range(3):
        byfile[f][k] += g[k]
for f, g in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print(f"  {f:32s} inst {g[0] / ti * 100:5.1f}%  lanes {g[1] / max(g[0], 1):5.1f}  samples {g[2] / ts * 100:5.1f}%")
print("-- top lines by samples")
for (f, l), g in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
    print(f"  {g[2] / ts * 100:5.1f}%s {g[0] / ti * 100:5.1f}%i lanes {g[1] / max(g[0], 1):5.1f}  {f}:{l}")
print("-- top lines by instructions")
for (f, l), g in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"  {g[0] / ti * 100:5.1f}%i {g[2] / ts * 100:5.1f}%s lanes {g[1] / max(g[0], 1):5.1f}  {f}:{l}")
print("-- stalls", {k: v for k, v in stall_tot.most_common(8)})
