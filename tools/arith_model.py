#!/usr/bin/env python
"""TEST INFRASTRUCTURE / design tool — error growth of candidate FAST arithmetics (tools/arith_model.c) against the oracle.

    python tools/arith_model.py Q dtype N ITERS variant [variant ...]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

SO = "/tmp/libarith_model.so"
subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", "-fPIC", "-shared", "-o", SO,
                       os.path.join(ROOT, "tools", "arith_model.c"), "-lm"])
M = C.CDLL(SO)
FLOOR = 1e-3
CHECK = (1, 2, 5, 10, 20, 50, 100, 150, 200, 300, 500, 1000)


def errors(a, ref, w, bulk):
    a64, r64 = a.astype(np.float64), ref.astype(np.float64)
    pe = pm = 0.0
    for k in range(ref.shape[0]):
        d = np.abs(a64[k] - r64[k])[bulk]
        r = np.abs(r64[k])[bulk]
        ok = r > FLOOR * w[k]
        pe = max(pe, float((d[ok] / r[ok]).max()))
        pm = max(pm, float(d.max() / r.max()))
    return pe, pm


def main():
    q, dt, n, iters = int(sys.argv[1]), np.dtype(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    variants = [int(v) for v in sys.argv[5:]]
    geom = int(os.environ.get("GEOM", "0"))
    O.set_threads(0)
    c, opp, w = O.tables(q)
    cls = O.classify(geom, n, n, n)
    mask = O.wall_mask(q, cls)
    omega = O.omega_cavity(n)
    pop = O.init_pop(q, cls, dt)
    bulk = cls == O.BULK
    ra, rb = pop.copy(), pop.copy()
    st = {v: [pop.copy(), pop.copy()] for v in variants}
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for t in range(1, iters + 1):
        O.step(q, ra, rb, cls, mask, omega)
        ra, rb = rb, ra
        for v in variants:
            a, b = st[v]
            M.model_step(C.c_int(v), C.c_int(dt == np.float64), C.c_int(q), p(c), p(opp), p(w), C.c_int(n), C.c_int(n), C.c_int(n),
                         p(a), p(b), p(cls), p(mask), C.c_double(omega))
            st[v] = [b, a]
        if t in CHECK or t == iters:
            line = f"it {t:5d}"
            for v in variants:
                pe, pm = errors(st[v][0], ra, w, bulk)
                line += f" | v{v}: {pe:.3e} {pm:.3e}"
            print(line, flush=True)


main()
