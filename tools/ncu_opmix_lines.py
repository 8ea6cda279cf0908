#!/usr/bin/env python
"""Executed warp instructions by source line (with the opcode mix of each line) for ONE kernel of an ncu report.
usage: ncu_opmix_lines.py REPORT.ncu-rep EXACT_MANGLED_KERNEL_SUBSTRING LIBRARY.so [N_LINES]
(the library must be the build the report was captured from: its -lineinfo line table maps SASS offsets to lines)"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kern, so = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1] if 'Source' in rows[1] else rows[0]
start0 = rows.index(hdr)+1
si, ii = hdr.index("Source"), hdr.index("Instructions Executed")
data=[]
for r in rows[start0:]:
    try: data.append((int(r[0],16), int(r[ii]), r[si].strip()))
    except Exception: pass
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    for cub in sorted(os.listdir(td)):
        if not cub.endswith(".cubin"): continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
        if kern in txt:
            sass = txt.split("\n"); break
start = [i for i, l in enumerate(sass) if l.startswith("\t.section\t.text.") and kern in l][0]
end = [i for i, l in enumerate(sass) if i > start and l.startswith("\t.section")][0]
cur, seq = None, {}
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur=(os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m: seq[int(m.group(1),16)] = cur
base=data[0][0]
tot=sum(d[1] for d in data)
per=collections.defaultdict(collections.Counter)
for addr,n,txt in data:
    t=txt.split(); op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    per[seq.get(addr-base)][op]+=n
src=open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'nav2_social_mpc_controller_b200', 'csrc', 'smpc_device.cuh')).read().split('\n')
lines=sorted(per.items(), key=lambda kv:-sum(kv[1].values()))
for k,c in lines[:int(sys.argv[4]) if len(sys.argv)>4 else 40]:
    s=sum(c.values())
    mix=' '.join('%s:%.1f'%(o,100*v/tot) for o,v in c.most_common(6))
    txt = src[k[1]-1].strip()[:70] if k and k[0]=='smpc_device.cuh' else ''
    print('%5.2f%% %s | %s | %s'%(100*s/tot, k, mix, txt))
