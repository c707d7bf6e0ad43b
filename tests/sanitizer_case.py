"""Small run of every kernel family, meant to be executed under compute-sanitizer (tests/test_gpu_sanitizer.py):
dense D3Q19 fp32 (vector width 4, x-face cache, wall fix-ups, REFERENCE and FAST arithmetic, the multi-iteration kernel, the
plain path of the REFERENCE evaluation), dense D3Q27 fp64, the TMA-fed variant, bGrid, set-up / rho-u kernels, and a
3-partition halo update on one device."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import neon_b200 as nb  # noqa: E402
from neon_b200 import problems as P  # noqa: E402


def main():
    bk = nb.Backend()
    skip_tma = os.environ.get("SANITIZER_SKIP_TMA") == "1"  # racecheck does not model the async (TMA) proxy
    for q, dt, dim, opts, arith in ((19, np.float32, (44, 18, 12), 0, nb.ARITH_FAST), (19, np.float32, (40, 10, 9), 0, nb.ARITH_REFERENCE),
                                    (27, np.float64, (36, 10, 9), 0, nb.ARITH_FAST), (19, np.float32, (64, 12, 8), nb.opt_kernel(nb.KERNEL_TMA), nb.ARITH_FAST)):
        if opts and skip_tma:
            continue
        grid = nb.dGrid(bk, dim)
        pop0, pop1, flag = P.setup_device(grid, q, dt, P.CAVITY_SPHERE)
        it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, 1.3, lattice_q=q, arith=arith,
                             opts=opts)
        for _ in range(3):
            it.run()
        if not opts:
            it.runMany(3)  # several iterations in one library call: the chain of dependent launches (plane-wise waits, coherent loads)
            it.runMany(2)
            every = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, 1.3, lattice_q=q, arith=arith,
                                    opts=15 << 16)  # NLBM_OPT_CHAIN_EARLY(15): every plane on the counters
            every.parity = it.parity
            every.runMany(3)
            it.parity = every.parity
        bk.syncAll()
        assert np.isfinite(it.getInput().updateHostData()).all()
    # REFERENCE arithmetic on a lid far outside the guard of its lean evaluation: fast and plain path inside the same warps
    grid = nb.dGrid(bk, (40, 12, 10))
    cls = P.host_classes(P.CAVITY_SPHERE, (40, 12, 10))
    pop0, pop1, flag = P.setup_host(grid, 19, np.float32, cls, P.host_populations(19, cls, np.float32, 0.4))
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, 1.1, arith=nb.ARITH_REFERENCE)
    for _ in range(4):
        it.run()
    bk.syncAll()
    # rho / u
    grid = nb.dGrid(bk, (24, 12, 10))
    pop0, pop1, flag = P.setup_device(grid, 19, np.float32, P.CAVITY)
    rho, u = grid.newField("rho", 1, np.float32), grid.newField("u", 3, np.float32)
    nb.LbmContainers.computeRhoAndU(pop0, flag, rho, u).run(0, nb.DataView.STANDARD)
    # three partitions on this device: views + halo pushes
    parts = []
    for p in range(3):
        g = nb.dGrid(bk, (36, 10, 12), partition=(p, 3))
        a, b, f = P.setup_device(g, 19, np.float32, P.CAVITY_SPHERE)
        parts.append((g, a, b, f))
    for g, a, b, f in parts:
        c = nb.LbmContainers.iteration(nb.StencilSemantic.streaming, a, b, f, 1.2)
        c.run(0, nb.DataView.INTERNAL)
        c.run(0, nb.DataView.BOUNDARY)
    # block-sparse
    bg = nb.bGrid(bk, (24, 16, 20))
    b0, b1, bf = P.setup_device(bg, 19, np.float32, P.CAVITY_SPHERE)
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, b0, b1, bf, 1.3)
    for _ in range(2):
        it.run()
    bk.syncAll()
    print("sanitizer case done")


if __name__ == "__main__":
    main()
