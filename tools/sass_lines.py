#!/usr/bin/env python
"""Print the SASS of a kernel whose inline chain passes through FILE lines [lo, hi] (nvdisasm -gi listing).
  tools/sass_lines.py disgi.txt kernel_substring file.cuh lo hi"""
import re
import sys

dis, kname, kfile, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
sec = False
chain = []
in_group = False
n = 0
for ln in open(dis):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m:
        sec = kname in m.group(1)
        continue
    if not sec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not in_group:
            chain = []
        in_group = True
        chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    in_group = False
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', ln)
    if m and any(f == kfile and lo <= l <= hi for f, l in chain):
        n += 1
        print(m.group(1), f"{chain[0][0][:16]}:{chain[0][1]}", m.group(2).strip()[:100])
print("instructions:", n, file=sys.stderr)
