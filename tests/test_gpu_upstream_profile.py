"""Long-run known answer from the reference repository itself: the u_x profile of the lid-driven cavity at N = 160 after 20 000
iterations that Autodesk/Neon ships as apps/lbmMultiRes/scripts/NeonUniformLBM_20000_Y.dat (written by the benchmark's --visual
mode, RunCavityTwoPop.cu:100-150; see tests/golden/upstream/README.md).  4.1 M cells x 20 000 iterations: hours on the reference's
CPU path, seconds here."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def nb():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    import neon_b200 as nb
    return nb


@pytest.fixture(scope="module")
def bk(nb):
    return nb.Backend()


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, ITERS, ULB = 160, 20000, 0.04


def profile(nb, bk, dtype, arith):
    from neon_b200 import problems as P
    grid = nb.dGrid(bk, (N, N, N))
    pop0, pop1, flag = P.setup_device(grid, 19, dtype, P.CAVITY, ulb=ULB)
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, nb.omega_from_re(N, 100.0, ULB),
                         arith=arith)
    done = 0
    while done < ITERS:  # CUDA-graph replay of the launch chain, 100 iterations per host call; the remainder one by one
        if ITERS - done >= 102:
            done += it.runGraph(100)
        else:
            it.run()
            done += 1
    assert done == ITERS
    rho, u = grid.newField("rho", 1, dtype), grid.newField("u", 3, dtype)
    nb.LbmContainers.computeRhoAndU(it.getInput(), flag, rho, u).run(0)
    bk.syncAll()
    return u.updateHostData()[0][N // 2, :, N // 2].astype(np.float64)  # u_x(y) at x = z = N/2  ([c][z][y][x])


@pytest.mark.parametrize("dtype,arith_name", [("float64", "reference"), ("float64", "fast"), ("float32", "reference"), ("float32", "fast")])
def test_profile_after_20000_iterations_matches_the_file_the_reference_ships(nb, bk, dtype, arith_name):
    ref = np.loadtxt(os.path.join(ROOT, "tests", "golden", "upstream", "NeonUniformLBM_20000_Y.dat"))
    assert ref.shape == (N, 2) and np.allclose(ref[:, 0], np.arange(N) / N, atol=1e-6)
    arith = nb.ARITH_REFERENCE if arith_name == "reference" else nb.ARITH_FAST
    ux = profile(nb, bk, np.dtype(dtype), arith)
    dev = np.abs(ux - ref[:, 1])
    rel = dev.max() / np.abs(ref[:, 1]).max()
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"upstream_profile_{dtype}_{arith_name}.json"), "w") as f:
        json.dump({"dtype": dtype, "arith": arith_name, "max_abs_dev": float(dev.max()), "max_rel_to_lid": float(rel), "argmax_y": int(dev.argmax()),
                   "ux": ux.tolist()}, f)
    assert ux[0] == 0.0 and abs(ux[N - 1] - ULB) < 1e-7  # wall and lid rows as the file has them
    if dtype == "float64":
        # the file prints 6 significant digits (operator<< of a double): every one of the 160 printed values is reproduced
        # digit for digit, in both arithmetics (measured on B200: largest deviation 0.4996 of the last printed digit)
        printed = np.array([float(f"{v:.6g}") for v in ux])
        assert np.array_equal(printed, ref[:, 1]), f"{int((printed != ref[:, 1]).sum())} of {N} printed values differ"
    else:
        # fp32 populations against the fp64 file: measured 6.4e-6 (FAST) / 6.1e-6 (REFERENCE) of the lid velocity
        assert rel < 2e-5, f"max deviation {dev.max():.3e} at y = {dev.argmax()} ({rel:.3e} of the lid velocity)"
