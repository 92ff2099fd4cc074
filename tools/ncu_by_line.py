#!/usr/bin/env python
"""Attribute the executed warp instructions of an ncu SASS page (ncu -i X.ncu-rep --page
source --csv --print-source sass) to source lines / functions using nvdisasm -g line info.
usage: ncu_by_line.py <sass.csv> <nvdisasm -g -c output> <kernel substring>"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, disasm, kern = sys.argv[1:4]
# address -> (file, line) for the kernel's section
addr_line = {}
cur = None
in_kernel = False
for ln in open(disasm, errors="replace"):
    if ln.startswith("//-") and ".text." in ln:
        in_kernel = kern in ln
        continue
    if not in_kernel:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        addr_line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
iA, iI, iT, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = None
per_line = defaultdict(lambda: [0, 0, 0])
tot = 0
for r in rows[2:]:
    try:
        a = int(r[iA], 16)
        n, t, s = int(r[iI]), int(r[iT]), int(r[iS])
    except (ValueError, IndexError):
        continue
    if base is None:
        base = a
    key = addr_line.get(a - base, ("?", 0))
    d = per_line[key]
    d[0] += n
    d[1] += t
    d[2] += s
    tot += n
print("total warp instructions", tot)
top = sorted(per_line.items(), key=lambda kv: -kv[1][0])
for (f, l), (n, t, s) in top[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print("%-18s %5d  %6.2f%%  lanes %5.1f  samples %6d" % (f, l, 100.0 * n / tot, t / max(n, 1), s))
