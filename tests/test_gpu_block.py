"""Parity of the CUDA block-sparse (bGrid) path with the CPU oracle — through the C ABI (nlbm_block_*).

The reference's bGrid with one partition reproduces its dGrid output byte for byte (SURVEY.md fact 4;
oracle/make_golden.py asserts it on the unmodified reference), so the dense oracle and the golden dumps are the checkers
here too.  Bars as for dGrid: flags / masks bit-exact, populations bit-exact in REFERENCE arithmetic, within 1e-5 / 1e-12
in FAST arithmetic.
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

REL_TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


@pytest.fixture(scope="module")
def nb():
    if not torch.cuda.is_available():
        pytest.fail("gpu tests need a CUDA device")
    import neon_b200 as nb
    return nb


@pytest.fixture(scope="module")
def bk(nb):
    return nb.Backend()


def rel_err(a, ref):
    q = ref.shape[0]
    scale = np.abs(ref.reshape(q, -1)).max(axis=1).reshape((q,) + (1,) * (ref.ndim - 1))
    return float((np.abs(a.astype(np.float64) - ref.astype(np.float64)) / scale).max())


def run_block(nb, bk, q, dtype, cls, pop, omega, iters, arith, active=None):
    from neon_b200 import problems as P
    nz, ny, nx = cls.shape
    grid = nb.bGrid(bk, (nx, ny, nz), active=active)
    pop0, pop1, flag = P.setup_host(grid, q, dtype, cls, pop)
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, omega, lattice_q=q,
                         arith=arith)
    for _ in range(iters):
        it.run()
    bk.syncAll()
    return it.getInput().updateHostData(), flag, grid


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4] for p in glob.glob(
    os.path.join(os.path.dirname(__file__), "golden", "*.npz"))))
def test_golden_reference_dumps_on_bgrid(nb, bk, golden_dir, name):
    from neon_b200 import problems as P
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    q = int(g["q"]) if "q" in g else 19
    cls, ref = g["cls"], g["pop"]
    pop = P.host_populations(q, cls, ref.dtype, float(g["ulb"]))
    out, flag, _ = run_block(nb, bk, q, ref.dtype, cls, pop, float(g["omega"]), int(g["iters"]), nb.ARITH_REFERENCE)
    assert np.array_equal(flag.masks(), g["mask"])
    assert np.array_equal(flag.classes(), cls)
    assert np.array_equal(out.view(np.uint8), ref.view(np.uint8))
    fast, _, _ = run_block(nb, bk, q, ref.dtype, cls, pop, float(g["omega"]), int(g["iters"]), nb.ARITH_FAST)
    assert rel_err(fast, ref) < REL_TOL[ref.dtype]


CASES = [
    # q, dtype, (nx, ny, nz), geom, iters
    (19, np.float32, (40, 24, 20), 1, 12),   # ragged in every direction: partially filled edge blocks
    (19, np.float32, (8, 8, 8), 0, 5),       # a single block
    (19, np.float32, (3, 3, 3), 0, 3),       # a single bulk cell
    (19, np.float64, (33, 17, 12), 1, 10),
    (27, np.float32, (40, 24, 20), 1, 8),
    (27, np.float64, (34, 18, 14), 1, 8),
    (19, np.float32, (48, 20, 24), 2, 10),   # flow over sphere: inlet as moving wall
    (27, np.float64, (48, 20, 24), 2, 6),
    (19, np.float32, (72, 64, 56), 1, 4),    # 9 x 8 x 7 blocks
]


@pytest.mark.parametrize("q,dtype,dim,geom,iters", CASES)
def test_parity_with_oracle(nb, bk, oracle, q, dtype, dim, geom, iters):
    nx, ny, nz = dim
    cls = oracle.classify(geom, nx, ny, nz)
    mask = oracle.wall_mask(q, cls)
    pop = oracle.init_pop(q, cls, dtype)
    omega = oracle.omega_cavity(max(dim))
    ref = oracle.run(q, pop, cls, mask, omega, iters)
    out, flag, _ = run_block(nb, bk, q, dtype, cls, pop, omega, iters, nb.ARITH_REFERENCE)
    assert np.array_equal(flag.masks(), mask), "wall masks must be bit-exact"
    assert np.array_equal(out.view(np.uint8), ref.view(np.uint8)), "REFERENCE arithmetic must be bit-exact"
    fast, _, _ = run_block(nb, bk, q, dtype, cls, pop, omega, iters, nb.ARITH_FAST)
    assert rel_err(fast, ref) < REL_TOL[np.dtype(dtype)]


def test_device_setup_matches_oracle(nb, bk, oracle):
    from neon_b200 import problems as P
    for geom in (0, 1, 2):
        for q, dt in ((19, np.float32), (27, np.float64)):
            nx, ny, nz = 37, 21, 18
            grid = nb.bGrid(bk, (nx, ny, nz))
            pop0, pop1, flag = P.setup_device(grid, q, dt, geom)
            bk.syncAll()
            cls = oracle.classify(geom, nx, ny, nz)
            assert np.array_equal(flag.classes(), cls)
            assert np.array_equal(flag.masks(), oracle.wall_mask(q, cls))
            ref = oracle.init_pop(q, cls, dt)
            assert np.array_equal(pop0.updateHostData().view(np.uint8), ref.view(np.uint8))
            assert np.array_equal(pop1.updateHostData().view(np.uint8), ref.view(np.uint8))


def test_sparse_grid_skips_solid_interior(nb, bk, oracle):
    """Blocks that lie entirely inside the solid sphere are not stored at all (the point of a bGrid); the fluid cells
    still match the dense oracle bit for bit, and cells that are not active stay untouched."""
    from neon_b200 import problems as P
    dim, q, iters = (64, 64, 64), 19, 6
    nx, ny, nz = dim
    sphere = (30.0, 33.0, 31.0, 22.0)
    cls = oracle.classify(1, nx, ny, nz, sphere)
    mask = oracle.wall_mask(q, cls)
    pop = oracle.init_pop(q, cls, np.float32)
    omega = 1.3
    ref = oracle.run(q, pop, cls, mask, omega, iters)
    cx, cy, cz, R = sphere
    active = lambda x, y, z: (x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2 >= (R - 2.0) ** 2  # keep a 2-cell solid shell
    grid = nb.bGrid(bk, dim, active=active)
    assert grid.n_blocks < 512, "some blocks must have been dropped"
    pop0, pop1, flag = P.setup_device(grid, q, np.float32, P.CAVITY_SPHERE, sphere)
    act = np.broadcast_to(active(*np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij", sparse=True)[::-1]), cls.shape)
    assert np.array_equal(flag.classes()[act], cls[act])
    assert (flag.classes()[~act] == nb.UNDEFINED).all()
    assert np.array_equal(flag.masks()[act], mask[act])
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, omega, arith=nb.ARITH_REFERENCE)
    for _ in range(iters):
        it.run()
    bk.syncAll()
    out = it.getInput().updateHostData()
    sel = np.broadcast_to(act, out.shape)
    assert np.array_equal(out[sel].view(np.uint8), ref[sel].view(np.uint8))
    assert not out[~sel].any()


def test_open_geometry_is_reported(nb, bk):
    grid = nb.bGrid(bk, (16, 8, 8))
    flag = grid.newFlagField()
    flag.setClasses(np.full((8, 8, 16), nb.BULK, np.int32))
    with pytest.raises(nb.NeonException):
        flag.computeWallNghMask(19)


@pytest.mark.parametrize("q,dtype,nparts", [(19, np.float32, 2), (27, np.float64, 2), (19, np.float32, 3)])
def test_partitions_on_one_gpu_match_single_partition(nb, bk, oracle, q, dtype, nparts):
    """Block-layer partitions + nlbm_block_halo_push (facing z-slice of the crossing populations) reproduce the
    single-partition result bit for bit, issued in OCC order: INTERNAL, halo, BOUNDARY.  (Upstream's bGrid halo is wrong
    for Q > 1 — SURVEY.md fact 4 — so the single-partition oracle is the reference.)"""
    from neon_b200 import _capi as capi
    from neon_b200 import problems as P
    dim, iters, omega = (40, 24, 8 * 2 * nparts + 5), 7, 1.25
    nx, ny, nz = dim
    cls = oracle.classify(1, nx, ny, nz)
    mask = oracle.wall_mask(q, cls)
    ref = oracle.run(q, oracle.init_pop(q, cls, dtype), cls, mask, omega, iters)
    parts = []
    for i in range(nparts):
        grid = nb.bGrid(bk, dim, partition=(i, nparts))
        pop0, pop1, flag = P.setup_device(grid, q, dtype, 1)
        parts.append((grid, [pop0, pop1], flag))
    eb = np.dtype(dtype).itemsize
    st = bk.streamHandle(0)
    for t in range(iters):
        a, b = t & 1, (t & 1) ^ 1
        conts = [nb.LbmContainers.iteration(nb.StencilSemantic.streaming, p[a], p[b], f, omega, q, arith=nb.ARITH_REFERENCE)
                 for _, p, f in parts]
        for c in conts:
            c.run(0, nb.DataView.INTERNAL)
        for i in range(nparts - 1):
            (g0, p0, _), (g1, p1, _) = parts[i], parts[i + 1]
            capi.call("nlbm_block_halo_push", C.byref(g0.desc(p0[a], None, None)), p0[a].data.data_ptr(),
                      C.byref(g1.desc(p1[a], None, None)), p1[a].data.data_ptr(), g1.n_blocks, eb, q, q, +1, st)
            capi.call("nlbm_block_halo_push", C.byref(g1.desc(p1[a], None, None)), p1[a].data.data_ptr(),
                      C.byref(g0.desc(p0[a], None, None)), p0[a].data.data_ptr(), g0.n_blocks + g0.n_ghost_down, eb, q, q, -1, st)
        for c in conts:
            c.run(0, nb.DataView.BOUNDARY)
    bk.syncAll()
    got = np.zeros_like(ref)
    for g, p, _ in parts:
        z0, z1 = g.layers[0] * 8, min(g.layers[1] * 8, nz)
        got[:, z0:z1] = p[iters & 1].updateHostData()[:, z0:z1]
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))


def test_plain_block_marker(nb, bk, oracle):
    """NLBM_FLAG_BLOCK_PLAIN (bit 30 of the flag word of a block's first cell): nlbm_block_wall_mask sets it exactly on the
    blocks whose 512 cells are all bulk without a wall neighbour, an upload of classes clears it, it never shows in the decoded
    classes / masks, and the step kernel computes the same bits whether it trusts the marker (default: such blocks load no
    flag words) or fetches every flag word (NLBM_OPT_FLAG_WORDS)."""
    from neon_b200 import problems as P
    from neon_b200 import _capi as capi
    nx, ny, nz = 72, 48, 40  # 9 x 6 x 5 blocks around a sphere
    cls = oracle.classify(1, nx, ny, nz)
    mask = oracle.wall_mask(19, cls)
    pop = oracle.init_pop(19, cls, np.float32)
    omega = oracle.omega_cavity(nx)
    grid = nb.bGrid(bk, (nx, ny, nz))
    outs = {}
    for opts in (0, capi.OPT_FLAG_WORDS):
        pop0, pop1, flag = P.setup_host(grid, 19, np.float32, cls, pop)
        raw = flag.cells[:grid.n_blocks].cpu().numpy().view(np.uint32)
        marked = (raw[:, 0] & 0x40000000) != 0
        assert not (raw[:, 1:] & 0x40000000).any(), "only the first cell of a block carries the marker"
        plain = ((raw & ~np.uint32(0x40000000)) == np.uint32(capi.BULK << capi.FLAG_CLASS_SHIFT)).all(axis=1)
        assert np.array_equal(marked, plain) and 0 < int(plain.sum()) < grid.n_blocks
        assert np.array_equal(flag.masks(), mask) and np.array_equal(flag.classes(), cls)
        for arith in (nb.ARITH_REFERENCE, nb.ARITH_FAST):
            a0, a1, _ = P.setup_host(grid, 19, np.float32, cls, pop)
            it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, a0, a1, flag, omega, arith=arith, opts=opts)
            for _ in range(6):
                it.run()
            bk.syncAll()
            outs[(opts, arith)] = it.getInput().updateHostData()
    ref = oracle.run(19, pop, cls, mask, omega, 6)
    for arith in (nb.ARITH_REFERENCE, nb.ARITH_FAST):
        assert np.array_equal(outs[(0, arith)].view(np.uint8), outs[(capi.OPT_FLAG_WORDS, arith)].view(np.uint8))
    assert np.array_equal(outs[(0, nb.ARITH_REFERENCE)].view(np.uint8), ref.view(np.uint8))
    # an upload of classes rewrites the words: no marker until the wall mask is built again
    flag.setClasses(cls)
    bk.syncAll()
    assert not (flag.cells[:grid.n_blocks].cpu().numpy().view(np.uint32) & 0x40000000).any()
