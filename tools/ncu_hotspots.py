#!/usr/bin/env python
"""Map an ncu report's per-SASS-instruction counters back to source lines (via nvdisasm -g line info)
and print opcode / source-line hot spots. Usage: tools/ncu_hotspots.py REPORT.ncu-rep KERNEL_SUBSTR [top_n]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "nav2_social_mpc_controller_b200", "libsmpc.so")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
si, ii, smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[0], 16), int(r[ii]), int(r[smp]), r[si].strip()))
    except Exception:
        pass
tot = sum(d[1] for d in data)
tots = max(1, sum(d[2] for d in data))
print(f"kernel rows {len(data)}  warp-instructions {tot}  stall samples {tots}")
ops, sm = collections.Counter(), collections.Counter()
for _, n, s, txt in data:
    parts = txt.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    ops[op] += n
    sm[op] += s
print("--- opcode mix")
for op, n in ops.most_common(18):
    print(f"{op:10s} {100*n/tot:5.1f}% inst  {100*sm[op]/tots:5.1f}% samples")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    sass = []
    for cub in sorted(os.listdir(td)):
        if not cub.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
        if kern in txt:
            sass = txt.split("\n")
            break
start = [i for i, l in enumerate(sass) if l.startswith("\t.section\t.text.") and kern in l][0]
end = [i for i, l in enumerate(sass) if i > start and l.startswith("\t.section")][0]
cur, seq = None, {}
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        seq[int(m.group(1), 16)] = cur
base = data[0][0]
agg, aggs = collections.Counter(), collections.Counter()
for addr, n, s, _ in data:
    key = seq.get(addr - base) or ("?", 0)
    agg[key] += n
    aggs[key] += s
src = {}
for fn in ("smpc_device.cuh", "smpc_kernels.cu", "smpc_kernels_nb.inc"):
    src[fn] = open(os.path.join(root, "nav2_social_mpc_controller_b200", "csrc", fn)).read().split("\n")
print("--- source lines")
for (fn, ln), n in agg.most_common(top):
    text = src[fn][ln - 1].strip()[:100] if fn in src and 0 < ln <= len(src[fn]) else ""
    print(f"{100*n/tot:5.1f}% inst {100*aggs[(fn, ln)]/tots:5.1f}% smp  {fn}:{ln}: {text}")
# coarse regions of smpc_device.cuh
def region(fn, ln):
    if fn != "smpc_device.cuh":
        return fn
    marks = [(i + 1, l) for i, l in enumerate(src[fn]) if l.startswith("__device__") or l.startswith("static __device__") or l.startswith("template <int NB>")]
    name = "?"
    for i, l in enumerate(src[fn]):
        if i + 1 > ln:
            break
        m = re.match(r"(?:static )?__device__ .*? (\w+)\(", l)
        if m:
            name = m.group(1)
    return name
reg = collections.Counter()
for (fn, ln), n in agg.items():
    reg[region(fn, ln)] += n
print("--- by enclosing function (line-info attribution)")
for k, n in reg.most_common(15):
    print(f"{100*n/tot:5.1f}%  {k}")
