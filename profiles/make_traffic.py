#!/usr/bin/env python
"""Reads `ncu --set full` reports and records the DRAM bytes per launch of the step kernels in profiles/traffic.json,
the file bench.py takes `roofline.traffic` from.

    python profiles/make_traffic.py KEY=report.ncu-rep [KEY=report.ncu-rep ...]
KEY is bench.py's workload key, e.g. d3q19_f32_512x512x512 or d3q19_f32_1024x512x512_bgrid.
"""
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def dram_bytes(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    out = {}
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(k)
        out[k] = float(r[i].replace(",", "")) * UNIT[units[i]]
    t = hdr.index("gpu__time_duration.sum")
    return {"kernel": r[hdr.index("Kernel Name")], "dram_bytes_read": out["dram__bytes_read.sum"],
            "dram_bytes_write": out["dram__bytes_write.sum"],
            "dram_bytes_per_launch": out["dram__bytes_read.sum"] + out["dram__bytes_write.sum"],
            "ncu_duration": f"{r[t]} {units[t]}", "report": os.path.basename(path)}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    for a in sys.argv[1:]:
        key, rep = a.split("=", 1)
        data[key] = dram_bytes(rep)
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
