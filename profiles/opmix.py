#!/usr/bin/env python
"""Executed-instruction mix of a kernel from an Nsight Compute report's source page (read here, no GPU needed).

    python profiles/opmix.py gpurun_out/prof.ncu-rep [cells_per_launch]
Prints, per SASS opcode, warp-level instructions executed (and per cell if the cell count is given), plus stall samples.
"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter


def main(path, cells=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    i_src, i_ex, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samp = Counter(), Counter()
    total = 0
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= i_ex:
            continue
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", r[i_src])
        if not m:
            continue
        op = m.group(2)
        key = op if op.startswith(("F2F", "IMAD.WIDE", "LDG", "STG", "LDL", "STL", "LDS", "LDGSTS", "MUFU")) else op.split(".")[0]
        n = int(r[i_ex] or 0)
        ops[key] += n
        samp[key] += int(r[i_samp] or 0)
        total += n
    print(f"total warp instructions executed: {total}" + (f"  = {total * 32 / cells:.1f} thread instructions per cell" if cells else ""))
    stot = sum(samp.values()) or 1
    for k, n in ops.most_common(45):
        per = f"{n * 32 / cells:8.1f}/cell" if cells else ""
        print(f"  {k:24s} {n:12d} {per}   stall samples {100.0 * samp[k] / stot:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
