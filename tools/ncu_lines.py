#!/usr/bin/env python
"""Per-source-line / per-opcode / per-stall summary of one kernel of an ncu report.
usage: ncu_lines.py <report.ncu-rep> <cubin-or-so> <kernel substring> [top N]
Needs the library compiled with -lineinfo (nvdisasm -g gives the address -> line map)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, binary, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
if binary.endswith(".so"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
else:
    cubins = [binary]
addr_line = {}
for cb in cubins:
    dis = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True, errors="replace").stdout
    cur, ink = None, False
    for ln in dis.splitlines():
        if ln.startswith("//-") and ".text." in ln:
            ink = kern in ln
            continue
        if not ink:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            addr_line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
H = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = None
line = collections.defaultdict(lambda: [0, 0, 0])
ops = collections.Counter()
opsamp = collections.Counter()
st = collections.Counter()
for r in rows[2:]:
    try:
        a = int(r[0], 16)
        n, t, s = int(r[H["Instructions Executed"]]), int(r[H["Thread Instructions Executed"]]), int(r[H["# Samples"]])
    except (ValueError, IndexError):
        continue
    if base is None:
        base = a
    d = line[addr_line.get(a - base, ("?", 0))]
    d[0] += n
    d[1] += t
    d[2] += s
    toks = r[1].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    ops[op] += n
    opsamp[op] += s
    for k in stalls:
        st[k] += int(r[H[k]] or 0)
tot = sum(v[0] for v in line.values())
samp = sum(v[2] for v in line.values())
print("warp instructions %d, samples %d" % (tot, samp))
print("by file:")
byf = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), v in line.items():
    for k in range(3):
        byf[f][k] += v[k]
for f, v in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print("  %-32s inst %5.1f%%  lanes %4.1f  samples %5.1f%%" % (f, 100.0 * v[0] / tot, v[1] / max(v[0], 1), 100.0 * v[2] / samp))
print("top lines:")
for (f, l), v in sorted(line.items(), key=lambda kv: -kv[1][0])[:top]:
    print("  %-18s %5d  inst %5.2f%%  lanes %4.1f  samples %5.2f%%" % (f, l, 100.0 * v[0] / tot, v[1] / max(v[0], 1), 100.0 * v[2] / samp))
print("opcodes:", ", ".join("%s %.1f%%" % (o, 100.0 * n / tot) for o, n in ops.most_common(16)))
print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(sum(st.values()), 1)) for k, v in st.most_common(9)))
# ---- by function of cw_core.cuh / phase of cw_kernels.cu (line ranges from the sources next to this script)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def ranges(path, pat):
    marks = []
    for i, ln in enumerate(open(path), 1):
        m = re.search(pat, ln)
        if m:
            marks.append((i, m.group(1)))
    return marks
def attribute(fname, marks):
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for (f, l), v in line.items():
        if f != fname:
            continue
        name = "(top)"
        for start, nm in marks:
            if l >= start:
                name = nm
        for k in range(3):
            agg[name][k] += v[k]
    return agg
core = os.path.join(root, "planner-rules-mcts_b200", "csrc", "cw_core.cuh")
kern_src = os.path.join(root, "planner-rules-mcts_b200", "csrc", "cw_kernels.cu")
print("by function (cw_core.cuh):")
for nm, v in sorted(attribute("cw_core.cuh", ranges(core, r"^BRM_HD\s+\S+\s+\**(\w+)\(")).items(), key=lambda kv: -kv[1][0]):
    print("  %-24s inst %5.1f%%  lanes %4.1f  samples %5.1f%%" % (nm, 100.0 * v[0] / tot, v[1] / max(v[0], 1), 100.0 * v[2] / samp))
print("by section (cw_kernels.cu):")
for nm, v in sorted(attribute("cw_kernels.cu", ranges(kern_src, r"^\s*// -{10,} (.*)$|^__device__ __forceinline__ \S+ (\w+)\(" )).items(), key=lambda kv: -kv[1][0]):
    print("  %-24s inst %5.1f%%  lanes %4.1f  samples %5.1f%%" % (nm, 100.0 * v[0] / tot, v[1] / max(v[0], 1), 100.0 * v[2] / samp))
