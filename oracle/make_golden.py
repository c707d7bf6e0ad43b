#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generate tests/golden/*.npz from the UNMODIFIED reference.

Runs oracle/_ref/ref_lbm (built by ``make -f oracle/Makefile.ref``: the
reference's own LbmIterationD3Q19 / LbmContainers on its CPU backend, driven by
oracle/ref_driver.cu) and stores the populations, wall masks and classes it
produced.  D3Q27 (``d3q27_*.npz``, key ``q`` = 27) comes from oracle/_ref/ref_lbm27
(``make -f oracle/Makefile.ref27``): the reference's apps/lbmMultiRes stream<T,27> /
collideBGK<T,27> containers and lattice tables, compiled unmodified with -DKBC, on a
one-level mGrid (oracle/ref_driver27.cu).  /root/reference is needed only to BUILD that binary; the fixtures
travel with the repo.  Also asserts on the spot that
  * the C oracle (oracle/lbm_oracle.c) reproduces every dump bit for bit,
  * bGrid (1 partition) and dGrid with 2 CPU partitions (--huGrid) give the
    same bytes as dGrid with 1 partition.

    python oracle/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_lbm")
OUT = os.path.join(ROOT, "tests", "golden")

# name, nx, ny, nz, iters, fp, geom
CASES = [
    ("cavity16_f32", 16, 16, 16, 20, "float", "cavity"),
    ("sphere16_f64", 16, 16, 16, 20, "double", "sphere"),
    ("sphere24_f32", 24, 24, 24, 30, "float", "sphere"),
    ("sphere20x12x16_f32", 20, 12, 16, 15, "float", "sphere"),
    ("cavity12_f64", 12, 12, 12, 40, "double", "cavity"),
]


# D3Q27: name, nx, ny, nz, iters, fp, geom (0 cavity, 1 cavity + sphere)
REF27 = os.path.join(ROOT, "oracle", "_ref", "ref_lbm27")
CASES27 = [
    ("d3q27_cavity12_f64", 12, 12, 12, 40, "double", 0),
    ("d3q27_sphere16_f64", 16, 16, 16, 20, "double", 1),
    ("d3q27_sphere20x12x16_f32", 20, 12, 16, 15, "float", 1),
    ("d3q27_cavity16_f32", 16, 16, 16, 20, "float", 0),
    ("d3q27_sphere24x20x28_f64", 24, 20, 28, 30, "double", 1),
    # omega of the dGrid benchmark (Config.cpp:105-111, nu from N - 2) instead of lidDrivenCavity.h's: what the C++ benchmark
    # app of this repo computes for --domain-size 16, so that the app can be checked against these bytes as well
    ("d3q27_sphere16_f64_cfgomega", 16, 16, 16, 25, "double", 1),
    ("d3q27_cavity16_f32_cfgomega", 16, 16, 16, 25, "float", 0),
]


def run_ref(tmp, nx, ny, nz, iters, fp, geom, grid="dGrid", ndev=1):
    path = os.path.join(tmp, f"d_{grid}_{ndev}.bin")
    subprocess.check_call(
        [REF, "--nx", str(nx), "--ny", str(ny), "--nz", str(nz), "--iters", str(iters), "--fp", fp, "--geom", geom,
         "--grid", grid, "--ndev", str(ndev), "--dump", path],
        cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return O.read_ref_dump(path)


def main():
    if not os.path.exists(REF):
        sys.exit("build the reference first: make -f oracle/Makefile.ref -j8")
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for name, nx, ny, nz, iters, fp, geom in CASES:
            d = run_ref(tmp, nx, ny, nz, iters, fp, geom)
            # the oracle must reproduce the reference bit for bit
            cls = O.classify(d["geom"], nx, ny, nz)
            mask = O.wall_mask(19, cls)
            assert np.array_equal(cls, d["cls"]) and np.array_equal(mask, d["mask"]), name
            pop = O.run(19, O.init_pop(19, cls, d["pop"].dtype), cls, mask, d["omega"], iters)
            assert np.array_equal(pop.view(np.uint8), d["pop"].view(np.uint8)), f"oracle != reference for {name}"
            # other reference paths that must agree with dGrid/1 partition
            if nx == ny == nz and nx % 8 == 0:
                b = run_ref(tmp, nx, ny, nz, iters, fp, geom, grid="bGrid")
                assert np.array_equal(b["pop"].view(np.uint8), d["pop"].view(np.uint8)), f"bGrid != dGrid for {name}"
            p2 = run_ref(tmp, nx, ny, nz, iters, fp, geom, ndev=2)
            assert np.array_equal(p2["pop"].view(np.uint8), d["pop"].view(np.uint8)), f"2 partitions != 1 for {name}"
            np.savez_compressed(os.path.join(OUT, name + ".npz"), pop=d["pop"], mask=d["mask"], cls=d["cls"],
                                omega=np.float64(d["omega"]), iters=np.int32(iters), geom=np.int32(d["geom"]),
                                ulb=np.float64(0.04))
            print(f"{name}: ok  ({d['pop'].nbytes} B populations, oracle bit-exact)")
        if not os.path.exists(REF27):
            sys.exit("build the D3Q27 reference driver: make -f oracle/Makefile.ref27 ref27")
        # the lattice tables the reference kernels index (lattice.h:15-77 under -DKBC) == the oracle's
        import json
        t = json.loads(subprocess.check_output([REF27, "--tables"], stderr=subprocess.DEVNULL))
        c, opp, w = O.tables(27)
        assert [r["c"] for r in t["rows"]] == c.tolist() and [r["opp"] for r in t["rows"]] == opp.tolist()
        assert np.array_equal(np.array([r["w"] for r in t["rows"]]), w)
        with open(os.path.join(OUT, "d3q27_tables.json"), "w") as f:  # as printed by the reference-compiled binary (%.17g)
            json.dump(t, f, indent=0)
        for name, nx, ny, nz, iters, fp, geom in CASES27:
            path = os.path.join(tmp, name + ".bin")
            om = ["--omega", repr(O.omega_cavity(nx))] if name.endswith("_cfgomega") else []
            subprocess.check_call([REF27, "--n", str(nx), str(ny), str(nz), "--iters", str(iters), "--fp", fp, "--geom", str(geom),
                                   "--dump", path] + om, cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            d = O.read_ref_dump(path)
            assert d["q"] == 27
            cls = O.classify(geom, nx, ny, nz)
            mask = O.wall_mask(27, cls)
            assert np.array_equal(cls, d["cls"]) and np.array_equal(mask, d["mask"]), name
            pop = O.run(27, O.init_pop(27, cls, d["pop"].dtype), cls, mask, d["omega"], iters)
            assert np.array_equal(pop.view(np.uint8), d["pop"].view(np.uint8)), f"oracle != reference for {name}"
            np.savez_compressed(os.path.join(OUT, name + ".npz"), pop=d["pop"], mask=d["mask"], cls=d["cls"],
                                omega=np.float64(d["omega"]), iters=np.int32(iters), geom=np.int32(geom), ulb=np.float64(0.04),
                                q=np.int32(27))
            print(f"{name}: ok  ({d['pop'].nbytes} B populations, oracle bit-exact)")


if __name__ == "__main__":
    main()
