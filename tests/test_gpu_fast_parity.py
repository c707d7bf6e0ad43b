"""NLBM_ARITH_FAST where it is benchmarked: long runs at the BASELINE.json configs' sizes against the CPU oracle.

north_star: "populations within 1e-5 relative (fp32) or 1e-12 (fp64) after N iterations".  The bench line runs FAST
arithmetic (fused multiply-add in the storage precision, reciprocal instead of three divisions), so the tolerance has to
hold over the iteration counts the benchmark uses — configs[0] is 100 iterations at 64^3 — not only over the 10-30
iterations of the small parity cases.  Two metrics, both must hold:

  per-element : max over bulk cells and populations of |a - ref| / |ref|, for |ref| above FLOOR x the population's lattice
                weight (in a cavity every bulk population stays within a few percent of its weight, so no element is excluded
                in practice; the floor only guards against a division by ~0);
  plane-max   : max |a - ref| / max|ref| per population (the looser metric of tests/test_gpu_dense.py).

The error-vs-iteration curve is written to gpurun_out/ (copied to profiles/ by the builder) so the growth can be judged.
REFERENCE arithmetic must stay bit-exact over the same runs.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}
FLOOR = 1e-3
CHECKPOINTS = (1, 2, 5, 10, 20, 50, 100, 150, 200, 300)


@pytest.fixture(scope="module")
def nb():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    import neon_b200 as nb
    return nb


def errors(a, ref, weights, bulk):
    a64, r64 = a.astype(np.float64), ref.astype(np.float64)
    per_elem, plane = 0.0, 0.0
    for k in range(ref.shape[0]):
        d = np.abs(a64[k] - r64[k])[bulk]
        r = np.abs(r64[k])[bulk]
        ok = r > FLOOR * weights[k]
        assert ok.mean() > 0.999, "the floor must not hide cells"
        per_elem = max(per_elem, float((d[ok] / r[ok]).max()))
        plane = max(plane, float(d.max() / r.max()))
    return per_elem, plane


@pytest.mark.parametrize("q,store,n,iters", [(19, np.float32, 64, 100), (19, np.float32, 128, 300), (27, np.float64, 64, 100),
                                             (19, np.float64, 64, 100), (27, np.float32, 64, 100)])
def test_fast_arithmetic_stays_within_tolerance_over_long_runs(nb, oracle, q, store, n, iters):
    from neon_b200 import problems as P
    oracle.set_threads(0)
    try:
        bk = nb.Backend()
        omega = nb.omega_from_re(n)  # Config.cpp:105-111: Re = 100, ulb = 0.04
        cls = oracle.classify(0, n, n, n)
        mask = oracle.wall_mask(q, cls)
        _, _, w = oracle.tables(q)
        bulk = cls == nb.BULK
        its = {}
        for arith in (nb.ARITH_FAST, nb.ARITH_REFERENCE):
            grid = nb.dGrid(bk, (n, n, n))
            pop0, pop1, flag = P.setup_device(grid, q, store, P.CAVITY)
            its[arith] = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, omega,
                                         lattice_q=q, arith=arith)
            assert np.array_equal(flag.masks(), mask)
        a = oracle.init_pop(q, cls, store)
        b = a.copy()
        curve = []
        for t in range(1, iters + 1):
            oracle.step(q, a, b, cls, mask, omega)
            a, b = b, a
            for it in its.values():
                it.run()
            if t in CHECKPOINTS or t == iters:
                bk.syncAll()
                exact = its[nb.ARITH_REFERENCE].getInput().updateHostData()
                assert np.array_equal(exact.view(np.uint8), a.view(np.uint8)), f"REFERENCE arithmetic differs at iteration {t}"
                pe, pm = errors(its[nb.ARITH_FAST].getInput().updateHostData(), a, w, bulk)
                curve.append({"iteration": t, "per_element": pe, "plane_max": pm})
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        name = f"fast_parity_d3q{q}_{np.dtype(store).name}_{n}.json"
        with open(os.path.join(out, name), "w") as f:
            json.dump({"lattice": f"D3Q{q}", "dtype": np.dtype(store).name, "box": n, "omega": omega, "tolerance": TOL[np.dtype(store)],
                       "floor": FLOOR, "curve": curve}, f, indent=1)
        worst = max(c["per_element"] for c in curve)
        assert worst < TOL[np.dtype(store)], curve
        assert max(c["plane_max"] for c in curve) < TOL[np.dtype(store)], curve
    finally:
        oracle.set_threads(1)
