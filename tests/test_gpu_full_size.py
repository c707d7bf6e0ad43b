"""Parity at the FULL sizes of BASELINE.json's configs, through a property of the scheme: an LBM iteration is local (a cell pulls
from its 18 / 26 neighbours), so after K iterations the state within R cells of a corner of the box depends only on the cells within
R + K of that corner.  The same corner of a SMALL box — which the CPU oracle iterates in a second — must therefore hold the same
bits (REFERENCE arithmetic) as the corner of the 512^3 ... 1024 x 1024 x 512 box on the GPU, as long as R + K stays clear of the
small box's far walls.  Both the corner at the origin and the one at (nx-1, ny-1, nz-1) — the lid, the outlet, and element offsets
beyond 2^31 and 2^33 — are compared, on dGrid and on bGrid."""
import gc

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

S, R, K = 48, 32, 6  # small box edge, compared region, iterations: R <= S - 1 - K
SPHERE = (392.0, 277.0, 256.0, 60.0)  # bench.py's flow-over-sphere workload (far from both corners)


@pytest.fixture(scope="module")
def nb():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    import neon_b200 as nb
    return nb


@pytest.fixture(scope="module")
def bk(nb):
    return nb.Backend()


def corner_of(pop, grid, dim, hi, is_block):
    """[q, R, R, R] of the low (hi = False) or high corner of the device field, without moving the rest of it"""
    nx, ny, nz = dim
    x0, y0, z0 = (nx - R, ny - R, nz - R) if hi else (0, 0, 0)
    if not is_block:
        return pop.view4[:, z0:z0 + R, y0:y0 + R, x0:x0 + R].cpu().numpy()
    bc = grid.block_coords[:grid.n_blocks]  # (bz, by, bx) of the local blocks
    sel = np.nonzero((bc[:, 0] >= z0 // 8) & (bc[:, 0] < (z0 + R) // 8) & (bc[:, 1] >= y0 // 8) & (bc[:, 1] < (y0 + R) // 8) &
                     (bc[:, 2] >= x0 // 8) & (bc[:, 2] < (x0 + R) // 8))[0]
    assert len(sel) == (R // 8) ** 3
    blocks = pop.view3[:, torch.from_numpy(sel).to(pop.view3.device)].cpu().numpy()  # [q, nsel, 512]
    out = np.empty((pop.cardinality, R, R, R), blocks.dtype)
    for i, b in enumerate(sel):
        bz, by, bx = bc[b][0] * 8 - z0, bc[b][1] * 8 - y0, bc[b][2] * 8 - x0
        out[:, bz:bz + 8, by:by + 8, bx:bx + 8] = blocks[:, i].reshape(-1, 8, 8, 8)
    return out


@pytest.mark.parametrize("q,dtype,dim,geom,is_block", [
    (19, np.float32, (512, 512, 512), 0, False),     # configs[1]
    (19, np.float32, (1024, 1024, 512), 0, False),   # half of configs[2]
    (19, np.float32, (1024, 1024, 1024), 0, False),  # configs[2] on ONE GPU: 2 x 81.6 GB of populations (skipped if that much is not free)
    (27, np.float64, (768, 768, 96), 0, False),      # configs[4], one GPU's slab
    (19, np.float32, (1024, 512, 512), 2, True),     # configs[3]: flow over a sphere on bGrid
])
def test_corners_of_full_size_boxes_match_the_oracle(nb, bk, oracle, q, dtype, dim, geom, is_block):
    from neon_b200 import problems as P
    assert R <= S - 1 - K and R % 8 == 0
    nx, ny, nz = dim
    omega = 1.3
    gc.collect()  # fields of earlier tests that only a reference cycle keeps alive
    torch.cuda.empty_cache()
    need = 2 * q * nx * ny * nz * np.dtype(dtype).itemsize + 5 * nx * ny * nz + (2 << 30)
    if torch.cuda.mem_get_info()[0] < need:
        pytest.skip(f"{need / 1e9:.0f} GB of device memory needed, {torch.cuda.mem_get_info()[0] / 1e9:.0f} GB free")
    grid = nb.bGrid(bk, dim) if is_block else nb.dGrid(bk, dim)
    sphere = SPHERE if geom == 2 else None
    outs = {}
    for arith in (nb.ARITH_REFERENCE, nb.ARITH_FAST):
        pop0, pop1, flag = P.setup_device(grid, q, dtype, geom, sphere)
        it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, omega, lattice_q=q, arith=arith)
        for _ in range(K):
            it.run()
        bk.syncAll()
        outs[arith] = [corner_of(it.getInput(), grid, dim, hi, is_block) for hi in (False, True)]
        del it, pop0, pop1, flag
        torch.cuda.empty_cache()
    for hi in (False, True):
        # the small box: [0, S)^3 of the big one, or its last S cells along every axis (sphere centre shifted with it)
        off = (nx - S, ny - S, nz - S) if hi else (0, 0, 0)
        sp = (SPHERE[0] - off[0], SPHERE[1] - off[1], SPHERE[2] - off[2], SPHERE[3]) if geom == 2 else None
        cls = oracle.classify(geom, S, S, S, sp)
        mask = oracle.wall_mask(q, cls)
        ref = oracle.run(q, oracle.init_pop(q, cls, dtype), cls, mask, omega, K)
        ref = ref[:, S - R:, S - R:, S - R:] if hi else ref[:, :R, :R, :R]
        got = outs[nb.ARITH_REFERENCE][1 if hi else 0]
        assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), f"REFERENCE arithmetic, {'high' if hi else 'low'} corner of {dim}"
        fast = outs[nb.ARITH_FAST][1 if hi else 0].astype(np.float64)
        scale = np.abs(ref.reshape(q, -1)).max(axis=1).reshape(q, 1, 1, 1)
        tol = 1e-5 if np.dtype(dtype) == np.float32 else 1e-12
        assert float((np.abs(fast - ref) / scale).max()) < tol
