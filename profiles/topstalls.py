#!/usr/bin/env python
"""Top stall sites of an ncu report (source page, SASS level).  python profiles/topstalls.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
start = 0
while start < len(rows):
    if rows[start] and rows[start][0] == "Kernel Name":
        print("kernel:", rows[start][1][:150])
        hdr = rows[start + 1]
        end = start + 2
        while end < len(rows) and not (rows[end] and rows[end][0] == "Kernel Name"):
            end += 1
        body = rows[start + 2:end]
        si = hdr.index("# Samples")
        tot = sum(int(r[si] or 0) for r in body)
        idx = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:n]
        stallcols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        print(f"total samples {tot}")
        for i in sorted(idx):
            r = body[i]
            st = sorted(((int(r[c] or 0), hdr[c]) for c in stallcols), reverse=True)[:2]
            print(f"{i:5d} {100*int(r[si])/max(tot,1):5.1f}%  {r[1].strip()[:70]:70s} {st}")
        start = end
    else:
        start += 1
