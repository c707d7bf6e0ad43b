#!/usr/bin/env python
"""Sweep driver for the lid-driven-cavity benchmark binary — the matrix the reference authors intended
(benchmarks/lbm-lid-driven-cavity-flow/lbm-lid-driven-cavity-flow.py:1-10,52-62: N = 64..512 step 64, grids, store/compute
precision pairs d/d, f/d, f/f, 1..n GPUs, warm-up 10, 100 iterations, 5 repetitions), with the same report-file naming so
upstream plotting picks the JSON files up unchanged.  Differences: cpu is not on the accelerated path and is skipped; eGrid
runs on the dense layout (the cavity activates every cell); --sOCC / transfer mode / halo semantic can be swept too.

    python sweep.py [--binary PATH] [--sizes 64 128 ...] [--gpus 8] [--grids dGrid bGrid] [--occ nOCC sOCC] [--out DIR] [--dry-run]
Writes one report JSON per configuration into --out plus sweep.csv (config columns + mean MLUPS).
"""
import argparse
import csv
import glob
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_BINARY = os.path.normpath(os.path.join(HERE, "..", "..", "bin", "lbm-lid-driven-cavity-flow"))


def configurations(a):
    device_sets = [" ".join(str(d) for d in range(n + 1)) for n in range(a.gpus)]
    for occ in a.occ:
        for n in a.sizes:
            for store, compute in (("double", "double"), ("float", "double"), ("float", "float")):
                for devs in device_sets:
                    for grid in a.grids:
                        if grid == "bGrid" and store != compute:
                            continue  # the block kernels exist for f/f and d/d
                        ndev = len(devs.split())
                        if grid == "bGrid" and ndev > 1 and (n + 7) // 8 // ndev < 2:
                            continue  # at least two block layers per device
                        if n < ndev:
                            continue
                        yield dict(occ=occ, n=n, store=store, compute=compute, devs=devs, grid=grid)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--binary", default=DEFAULT_BINARY)
    ap.add_argument("--sizes", nargs="+", type=int, default=[64, 128, 192, 256, 320, 384, 448, 512])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--grids", nargs="+", default=["dGrid", "bGrid", "eGrid"])
    ap.add_argument("--occ", nargs="+", default=["nOCC"], choices=["nOCC", "sOCC"])
    ap.add_argument("--transfer", default="get", choices=["get", "put"])
    ap.add_argument("--semantic", default="huLattice", choices=["huLattice", "huGrid"])
    ap.add_argument("--warmup-iter", type=int, default=10)
    ap.add_argument("--max-iter", type=int, default=100)
    ap.add_argument("--repetitions", type=int, default=5)
    ap.add_argument("--out", default="sweep_out")
    ap.add_argument("--device-setup", action="store_true", help="set the problem up on the device (no host mirrors)")
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args()

    cfgs = list(configurations(a))
    if a.dry_run:
        for c in cfgs:
            print(c)
        print(len(cfgs), "configurations")
        return 0
    os.makedirs(a.out, exist_ok=True)
    rows = []
    with open(os.path.join(a.out, "sweep.log"), "w") as log:
        for i, c in enumerate(cfgs):
            name = (f"lbm-lid-driven-cavity-flow___gpu_{c['n']}_{c['store']}_{c['compute']}_{c['devs'].replace(' ', '_')}_{c['occ']}"
                    f"_{c['grid']}")
            cmd = [a.binary, "--deviceType", "gpu", "--deviceIds", *c["devs"].split(), "--grid", c["grid"], "--domain-size", str(c["n"]),
                   "--warmup-iter", str(a.warmup_iter), "--repetitions", str(a.repetitions), "--max-iter", str(a.max_iter),
                   "--report-filename", os.path.join(a.out, name), "--computeFP", c["compute"], "--storageFP", c["store"], "--benchmark",
                   "--" + c["occ"], "--" + a.transfer, "--" + a.semantic] + (["--device-setup"] if a.device_setup else [])
            log.write("\n-------------------------------------------\n" + " ".join(cmd) + "\n-------------------------------------------\n")
            log.flush()
            r = subprocess.run(cmd, text=True, stdout=log, stderr=subprocess.STDOUT)
            mlups = None
            reports = sorted(glob.glob(os.path.join(a.out, name + "_*.json")))
            if r.returncode == 0 and reports:
                vals = json.load(open(reports[-1]))["MLUPS"]
                mlups = sum(vals) / len(vals)
            rows.append(dict(c, mlups=mlups, rc=r.returncode))
            sys.stdout.write(f"\r[{i + 1}/{len(cfgs)}] {name}: {mlups}      ")
            sys.stdout.flush()
    with open(os.path.join(a.out, "sweep.csv"), "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()) if rows else ["n"])
        w.writeheader()
        w.writerows(rows)
    print(f"\n{len(rows)} runs, {sum(1 for r in rows if r['rc'] != 0)} failed -> {a.out}/sweep.csv")
    return 0


if __name__ == "__main__":
    sys.exit(main())
