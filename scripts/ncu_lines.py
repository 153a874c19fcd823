#!/usr/bin/env python
"""Summarise an ncu report per source line: share of executed instructions / stall samples + top stall reasons.
usage: ncu_lines.py report.ncu-rep [min_pct]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
funcs = collections.OrderedDict()
cur = hdr = fname = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur = funcs.setdefault(r[1].split("(")[0][-40:], collections.OrderedDict())
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if cur is None or hdr is None or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
    sc = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    key = (fname, ln, r[1].strip()[:64])
    e = cur.setdefault(key, [0, 0, collections.Counter()])
    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    e[0] += num(r[iI])
    e[1] += num(r[iS])
    for i in sc:
        e[2][hdr[i][6:]] += num(r[i])
for fn, agg in funcs.items():
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"===== {fn}  inst={ti} samples={ts}")
    for (f, ln, src), (i, s, st) in agg.items():
        if 100 * s / ts > thr or 100 * i / ti > thr:
            print(f"{f[:14]:14s}{ln:4d} inst {100*i/ti:5.1f}% smp {100*s/ts:5.1f}%  {src:64s} {st.most_common(2)}")
