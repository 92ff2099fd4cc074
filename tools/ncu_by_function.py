#!/usr/bin/env python
"""Share of executed warp instructions per device function (uses tools/ncu_by_line.py output
and the function start lines of the given .cuh/.cu sources).
usage: ncu_by_function.py <byline.txt> <source files...>"""
import re
import sys

lines = open(sys.argv[1]).read().splitlines()[1:]
funcs = {}
for path in sys.argv[2:]:
    name = path.split("/")[-1]
    src = open(path).read().splitlines()
    fl = []
    for i, l in enumerate(src, 1):
        m = re.match(r"\s*(?:template.*>\s*)?(?:__device__|__global__).*?\b([A-Za-z_]\w+)\(", l)
        if m:
            fl.append((i, m.group(1)))
        elif re.match(r"\s*(__device__|__global__)", l) and i < len(src):
            m2 = re.search(r"\b([A-Za-z_]\w+)\(", src[i])
            if m2:
                fl.append((i, m2.group(1)))
    funcs[name] = sorted(fl)


def func_of(f, n):
    name = "-"
    for s, fn in funcs.get(f, []):
        if s <= n:
            name = fn
    return name


agg = {}
for l in lines:
    p = l.split()
    if len(p) < 7:
        continue
    f, ln, pct, lanes = p[0], int(p[1]), float(p[2].rstrip("%")), float(p[4])
    key = (f, func_of(f, ln))
    d = agg.setdefault(key, [0.0, 0.0])
    d[0] += pct
    d[1] += pct * lanes
for k, (pct, pl) in sorted(agg.items(), key=lambda x: -x[1][0])[:20]:
    if pct > 0:
        print("%-22s %-24s %6.2f%%  lanes %5.1f" % (k[0], k[1], pct, pl / pct))
