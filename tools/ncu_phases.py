#!/usr/bin/env python
"""Attribute an ncu source-page CSV to the OUTERMOST source line of the kernel file (through the inline chains of
`nvdisasm -gi`), so that the cost of inlined helpers lands on the phase of the kernel that called them.

  ncu -i rep.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all lib.so; nvdisasm -gi -c lib.sm_100a.cubin > disgi.txt
  tools/ncu_phases.py src.csv disgi.txt kernel_substring kernel_file.cuh [line:label ...]

Optional `line:label` arguments cut the kernel file into phases (a phase starts at its line)."""
import collections
import csv
import re
import sys

src_csv, dis_txt, kname, kfile = sys.argv[1:5]
cuts = sorted((int(a.split(':')[0]), a.split(':', 1)[1]) for a in sys.argv[5:])
loc_of = {}
chain = []       # the current location group: [(file, line), ...] innermost first
in_group = False  # the previous line was a location marker
infn = False
label = None
for ln in open(dis_txt):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m:
        infn = kname in m.group(1)
        chain = []
        label = None
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not in_group:
            chain = []
        in_group = True
        chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    in_group = False
    m = re.match(r'\s*(\$\S+):\s*$', ln)
    if m:
        label = m.group(1)
        chain = []
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        inner = chain[0] if chain else ('?', 0)
        outer = None
        for f, l in chain:
            if f == kfile:
                outer = l
        loc_of[int(m.group(1), 16)] = (inner, outer, label)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, iinst, ithr, isamp = (hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"),
                          hdr.index("# Samples"))
isrc = hdr.index("Source")
data = [r for r in rows[2:] if len(r) > isamp]


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def addr(x):
    return int(x, 16) if x.lower().startswith('0x') else int(x)


base = addr(data[0][ia])
by_line = collections.defaultdict(lambda: [0, 0, 0])
by_fn = collections.defaultdict(lambda: [0, 0, 0])
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
phase_stalls = collections.defaultdict(collections.Counter)
ops = collections.defaultdict(collections.Counter)


def phase_of(line):
    name = 'pre'
    for l, nm in cuts:
        if line >= l:
            name = nm
    return name


for r in data:
    inner, outer, lab = loc_of.get(addr(r[ia]) - base, (('?', 0), None, None))
    key = outer if outer is not None else f"{lab or '?'}"
    g = by_line[key]
    g[0] += num(r[iinst]); g[1] += num(r[ithr]); g[2] += num(r[isamp])
    ph = phase_of(outer) if outer is not None else f"call {lab}"
    g = by_fn[ph]
    g[0] += num(r[iinst]); g[1] += num(r[ithr]); g[2] += num(r[isamp])
    for i in stall_cols:
        phase_stalls[ph][hdr[i]] += num(r[i])
    op = r[isrc].split()[0] if r[isrc].split() else '?'
    if op.startswith('@'):
        op = r[isrc].split()[1] if len(r[isrc].split()) > 1 else op
    ops[ph][op.split('.')[0]] += num(r[iinst])
ti = sum(g[0] for g in by_line.values())
ts = sum(g[2] for g in by_line.values())
print(f"total warp-inst {ti} samples {ts}")
print("-- phases")
for ph, g in sorted(by_fn.items(), key=lambda x: -x[1][2]):
    st = ' '.join(f"{k[6:]}={v / max(g[2], 1) * 100:.0f}%" for k, v in phase_stalls[ph].most_common(4))
    print(f"  {ph:28s} inst {g[0] / ti * 100:5.1f}%  lanes {g[1] / max(g[0], 1):5.1f}  samples {g[2] / ts * 100:5.1f}%   {st}")
    print("       ops: " + ' '.join(f"{k}={v / max(g[0], 1) * 100:.0f}%" for k, v in ops[ph].most_common(10)))
print("-- kernel lines by samples")
for key, g in sorted(by_line.items(), key=lambda x: -x[1][2])[:45]:
    print(f"  {g[2] / ts * 100:5.1f}%s {g[0] / ti * 100:5.1f}%i lanes {g[1] / max(g[0], 1):5.1f}  {key}")
