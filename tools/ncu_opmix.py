#!/usr/bin/env python
"""Executed warp instructions by opcode from an ncu report captured with --import-source on.
usage: ncu_opmix.py REPORT.ncu-rep"""
import collections, csv, io, subprocess, sys
rep=sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1] if 'Source' in rows[1] else rows[0]
start = rows.index(hdr)+1
si, ii = hdr.index("Source"), hdr.index("Instructions Executed")
tot=0; ops=collections.Counter()
data=[]
for r in rows[start:]:
    try:
        n=int(r[ii]); txt=r[si].strip()
    except Exception: continue
    t=txt.split()
    op=t[1] if t[0].startswith('@') else t[0]
    data.append((n,op,txt))
    ops[op.split('.')[0]]+=n; tot+=n
print('total',tot)
for k,v in ops.most_common(40): print('%-10s %6.2f%%'%(k,100*v/tot))
