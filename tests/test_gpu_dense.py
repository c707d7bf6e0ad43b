"""Parity of the CUDA dense (dGrid) path with the CPU oracle — through the C ABI (libneon_lbm.so).

Bars (BASELINE.json north_star): flags and wall masks bit-exact; populations bit-exact in NLBM_ARITH_REFERENCE mode,
within 1e-5 (fp32) / 1e-12 (fp64) relative in NLBM_ARITH_FAST mode.
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
    """max |a - ref| relative to the largest reference magnitude of each population"""
    q = ref.shape[0]
    scale = np.abs(ref.reshape(q, -1)).max(axis=1).reshape((q,) + (1,) * (ref.ndim - 1))
    return float((np.abs(a.astype(np.float64) - ref.astype(np.float64)) / scale).max())


def run_cuda(nb, bk, q, dtype, cls, pop, omega, iters, arith, compute=None, opts=0, occ=None):
    from neon_b200 import problems as P
    nz, ny, nx = cls.shape
    grid = nb.dGrid(bk, (nx, ny, nz))
    pop0, pop1, flag = P.setup_host(grid, q, dtype, cls, pop)
    it = nb.LbmIteration(nb.StencilSemantic.streaming, occ or nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, omega,
                         lattice_q=q, compute=compute, arith=arith, opts=opts)
    for _ in range(iters):
        it.run()
    bk.syncAll()
    return it.getInput().updateHostData(), flag


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4] for p in glob.glob(
    os.path.join(os.path.dirname(__file__), "golden", "*.npz"))))
def test_golden_reference_dumps(nb, bk, golden_dir, name):
    """Populations / masks dumped by the UNMODIFIED reference (oracle/make_golden.py): REFERENCE mode is bit-exact."""
    from neon_b200 import problems as P
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    q = int(g["q"]) if "q" in g else 19  # d3q27_*: the reference's apps/lbmMultiRes kernels (oracle/ref_driver27.cu)
    cls, ref = g["cls"], g["pop"]
    pop = P.host_populations(q, cls, ref.dtype, float(g["ulb"]))
    for kern in (nb.KERNEL_DIRECT, nb.KERNEL_TMA):
        out, flag = run_cuda(nb, bk, q, ref.dtype, cls, pop, float(g["omega"]), int(g["iters"]), nb.ARITH_REFERENCE,
                             opts=nb.opt_kernel(kern))
        assert np.array_equal(flag.masks(), g["mask"])
        assert np.array_equal(flag.classes(), cls)
        assert np.array_equal(out.view(np.uint8), ref.view(np.uint8)), f"kernel {kern}"
    fast, _ = run_cuda(nb, bk, q, ref.dtype, cls, pop, float(g["omega"]), int(g["iters"]), nb.ARITH_FAST)
    assert rel_err(fast, ref) < REL_TOL[ref.dtype]


CASES = [
    # q, store, compute, (nx, ny, nz), geom, iters
    (19, np.float32, None, (40, 24, 20), 1, 12),
    (19, np.float32, None, (130, 9, 7), 0, 6),      # nx beyond one warp segment, ragged tail
    (19, np.float32, None, (3, 3, 3), 0, 3),        # a single bulk cell
    (19, np.float32, None, (300, 7, 5), 0, 4),      # wider than one TMA tile, ragged in x and y
    (19, np.float64, None, (270, 5, 6), 1, 4),
    (19, np.float32, None, (300, 70, 40), 1, 3),    # ~2000 TMA tiles: every CTA's stage ring wraps several times
    (27, np.float64, None, (300, 40, 44), 1, 3),
    (27, np.float32, None, (200, 50, 40), 2, 3),
    (19, np.float64, None, (260, 48, 40), 2, 3),
    (19, np.float64, None, (33, 17, 12), 1, 10),
    (19, np.float32, np.float64, (36, 20, 16), 1, 8),
    (27, np.float32, None, (40, 24, 20), 1, 10),
    (27, np.float64, None, (34, 18, 14), 1, 10),
    (19, np.float32, None, (48, 20, 24), 2, 10),    # flow over sphere: inlet as moving wall
    (27, np.float64, None, (48, 20, 24), 2, 6),
]


@pytest.mark.parametrize("q,store,compute,dim,geom,iters", CASES)
def test_parity_with_oracle(nb, bk, oracle, q, store, compute, dim, geom, iters):
    nx, ny, nz = dim
    cls = oracle.classify(geom, nx, ny, nz)
    mask = oracle.wall_mask(q, cls)
    pop = oracle.init_pop(q, cls, store)
    omega = oracle.omega_cavity(max(dim))
    ref = oracle.run(q, pop, cls, mask, omega, iters, compute)
    for kern in (nb.KERNEL_TMA, nb.KERNEL_DIRECT):  # both kernels: same bits as the reference
        out, flag = run_cuda(nb, bk, q, store, cls, pop, omega, iters, nb.ARITH_REFERENCE, compute, opts=nb.opt_kernel(kern))
        assert np.array_equal(flag.masks(), mask), "wall masks must be bit-exact"
        assert np.array_equal(out.view(np.uint8), ref.view(np.uint8)), f"REFERENCE arithmetic must be bit-exact (kernel {kern})"
        fast, _ = run_cuda(nb, bk, q, store, cls, pop, omega, iters, nb.ARITH_FAST, compute, opts=nb.opt_kernel(kern))
        assert rel_err(fast, ref) < REL_TOL[np.dtype(store)]
    # every vector width of the direct kernel computes the same bits (the default is 2 cells per thread for views of up to
    # 2^20 cells and 4 above: small boxes would otherwise never run the 16-byte path)
    for vec in (1, 2, 4):
        v, _ = run_cuda(nb, bk, q, store, cls, pop, omega, iters, nb.ARITH_REFERENCE, compute,
                        opts=nb.opt_vec(vec) | nb.opt_kernel(nb.KERNEL_DIRECT))
        assert np.array_equal(v.view(np.uint8), ref.view(np.uint8)), f"vec={vec}"


def _cluttered_cavity(oracle, dim, seed, density=0.08):
    """Cavity with random bounce-back / moving-wall cells in the interior, many of them right next to the x faces: the
    cells at x = 1 and x = nx-2 carry every kind of wall-bit pattern, so the speculative x-face fix-up of the direct kernel
    is taken, refused and mixed with the late path within one warp."""
    nx, ny, nz = dim
    cls = oracle.classify(0, nx, ny, nz)
    rng = np.random.default_rng(seed)
    inner = cls[1:-1, 1:-1, 1:-1]
    r = rng.random(inner.shape)
    inner[r < density] = 0        # bounce-back obstacle cells
    inner[r > 1.0 - 0.02] = 1     # a few moving-wall cells (non-zero wall populations)
    col = rng.random((nz - 2, ny - 2)) < 0.3
    inner[:, :, 1][col] = 0       # x = 2 solid: the cell at x = 1 then has walls on both sides along x
    inner[:, :, -2][col.T[: ny - 2, : nz - 2].T] = 0
    return cls


VARIANTS = [0, "OPT_REF_LITERAL", "OPT_FLAG_WORDS", "OPT_FLAGS_SUMMARY_FIRST", "OPT_NO_XFACE_FIXUP_PREFETCH", "OPT_NO_XFACE_PREFETCH",
            ("OPT_FLAG_WORDS", "OPT_NO_XFACE_FIXUP_PREFETCH", "OPT_NO_XFACE_PREFETCH")]


@pytest.mark.parametrize("q,store,dim", [(19, np.float32, (136, 22, 18)), (19, np.float32, (64, 16, 12)), (27, np.float64, (70, 14, 12)),
                                         (19, np.float64, (37, 12, 11)), (27, np.float32, (131, 10, 9))])
def test_kernel_variants_agree_bit_for_bit(nb, bk, oracle, q, store, dim):
    """Every way the direct kernel learns about its cells (cell map / flag words / row summary) and every speculative fetch
    (kept wall values, x-face fix-up operands) switched on or off: the same bits as the oracle, on the cavity, on the inlet
    geometry (non-zero populations in the x = 0 wall) and on a cavity cluttered with obstacles next to the x faces."""
    for geom in (0, 2, "clutter"):
        nx, ny, nz = dim
        cls = _cluttered_cavity(oracle, dim, 7) if geom == "clutter" else oracle.classify(geom, nx, ny, nz)
        mask = oracle.wall_mask(q, cls)
        pop = oracle.init_pop(q, cls, store)
        # wall populations that differ from cell to cell: a wrong operand cannot cancel
        rng = np.random.default_rng(3)
        noise = (rng.random(pop.shape) * 0.01).astype(store)
        pop = np.where(np.broadcast_to(cls != nb.BULK, pop.shape), pop + noise, pop).astype(store)
        omega, iters = 1.3, 5
        ref = oracle.run(q, pop, cls, mask, omega, iters)
        for var in VARIANTS:
            names = var if isinstance(var, tuple) else ((var,) if var else ())
            opts = nb.opt_kernel(nb.KERNEL_DIRECT)
            for n in names:
                opts |= getattr(nb, n)
            out, flag = run_cuda(nb, bk, q, store, cls, pop, omega, iters, nb.ARITH_REFERENCE, opts=opts)
            assert np.array_equal(flag.masks(), mask)
            assert np.array_equal(out.view(np.uint8), ref.view(np.uint8)), f"geom {geom} variant {var}"


def test_exact_building_blocks(nb):
    """The two pieces of the conversion-lean REFERENCE evaluation that are not plain IEEE operations, checked on the device
    against the instructions they replace: float -> double through an integer multiply-add for every positive normal float
    (2^31 - 2^24 bit patterns), and the shared-reciprocal division on 3 x 2^30 random quotients inside its guard."""
    import ctypes as C
    from neon_b200 import _capi as capi
    bad = C.c_uint64(123)
    capi.call("nlbm_selftest_exact", 0, 0, 0, C.byref(bad))
    assert bad.value == 0, f"{bad.value} floats widen differently"
    for seed in (1, 2):
        capi.call("nlbm_selftest_exact", 1, 1 << 29, seed, C.byref(bad))
        assert bad.value == 0, f"{bad.value} quotients differ from IEEE division"


@pytest.mark.parametrize("ulb", [0.2, 0.45])
def test_reference_bits_outside_the_guard_of_the_lean_evaluation(nb, bk, oracle, ulb):
    """REFERENCE arithmetic for D3Q19 fp32 runs the conversion-lean evaluation (csrc/lbm_collide_exact.cuh) where its guard
    holds (positive populations, |u| < 0.1) and the plain one elsewhere, cell by cell.  A lid far faster than any valid
    simulation (Mach > 0.3, negative populations next to the lid) makes both paths run inside the same warps: the bits
    must still be the oracle's, on dGrid with both kernels and on bGrid."""
    from neon_b200 import problems as P
    dim = (72, 24, 16)
    cls = oracle.classify(1, *dim)
    mask = oracle.wall_mask(19, cls)
    pop = oracle.init_pop(19, cls, np.float32, ulb)
    assert np.array_equal(pop, P.host_populations(19, cls, np.float32, ulb))
    omega, iters = 1.1, 8
    ref = oracle.run(19, pop, cls, mask, omega, iters)
    assert np.isfinite(ref).all()
    rho_u = oracle.rho_u(ref, cls, mask)
    assert np.abs(rho_u[1][:, cls == oracle.BULK]).max() > 0.12, "the case must leave the guard"
    for opts in (nb.opt_kernel(nb.KERNEL_DIRECT), nb.opt_kernel(nb.KERNEL_TMA), nb.opt_kernel(nb.KERNEL_DIRECT) | nb.OPT_REF_LITERAL):
        out, _ = run_cuda(nb, bk, 19, np.float32, cls, pop, omega, iters, nb.ARITH_REFERENCE, opts=opts)
        assert np.array_equal(out.view(np.uint8), ref.view(np.uint8)), hex(opts)
    grid = nb.bGrid(bk, dim)
    pop0, pop1, flag = P.setup_host(grid, 19, np.float32, cls, pop)
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, omega, arith=nb.ARITH_REFERENCE)
    for _ in range(iters):
        it.run()
    bk.syncAll()
    assert np.array_equal(it.getInput().updateHostData().view(np.uint8), ref.view(np.uint8))


@pytest.mark.parametrize("q,store,dim,geom,iters", [(19, np.float32, (40, 24, 20), 1, 7), (19, np.float32, (64, 64, 64), 0, 12), (27, np.float64, (34, 18, 14), 1, 5),
                                                    (19, np.float64, (33, 17, 12), 2, 6), (27, np.float32, (70, 20, 24), 1, 4),
                                                    (19, np.float32, (300, 40, 30), 1, 3)])
def test_several_iterations_in_one_launch(nb, bk, oracle, q, store, dim, geom, iters):
    """nlbm_dense_step_n (LbmIteration.runMany): a chain of dependent launches, iteration t+1 launched while t still runs, the
    tiles of the first planes waiting plane-wise for the tiles of the previous iteration they depend on, the others for the end of
    the previous launch; the field the previous iteration wrote is read through coherent loads.  Same
    bits as the oracle in REFERENCE arithmetic, the same bits as iteration-by-iteration launches in FAST arithmetic, odd and even
    iteration counts (where the result lands), more tiles than resident blocks (300 x 40 x 30) and fewer, every vector width, and
    the chain captured into a CUDA graph and replayed."""
    from neon_b200 import problems as P
    nx, ny, nz = dim
    cls = oracle.classify(geom, nx, ny, nz)
    mask = oracle.wall_mask(q, cls)
    pop = oracle.init_pop(q, cls, store)
    omega = oracle.omega_cavity(max(dim))
    ref = oracle.run(q, pop, cls, mask, omega, iters + 1)
    one = None
    for arith in (nb.ARITH_REFERENCE, nb.ARITH_FAST):
        for mode in (0, nb.opt_vec(1), nb.opt_vec(2), nb.opt_vec(4), 15 << 16, 2 << 16):  # (NLBM_OPT_CHAIN_EARLY: all planes / two planes on the counters)
            grid = nb.dGrid(bk, dim)
            pop0, pop1, flag = P.setup_host(grid, q, store, cls, pop)
            it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, omega, lattice_q=q, arith=arith,
                                 opts=mode)
            it.runMany(iters)       # one library call
            it.runMany(1)           # and one more through the same entry point: parity bookkeeping
            bk.syncAll()
            many = it.getInput().updateHostData()
            if arith == nb.ARITH_REFERENCE:
                assert np.array_equal(many.view(np.uint8), ref.view(np.uint8)), f"mode {mode:#x}"
            else:
                if one is None:
                    one, _ = run_cuda(nb, bk, q, store, cls, pop, omega, iters + 1, nb.ARITH_FAST)
                assert np.array_equal(many.view(np.uint8), one.view(np.uint8)), f"mode {mode:#x}"
            assert np.array_equal(flag.masks(), mask)
    # the chain inside a CUDA graph, replayed: 2 plain + 3 x n iterations
    n = iters + (iters & 1)
    grid = nb.dGrid(bk, dim)
    pop0, pop1, flag = P.setup_host(grid, q, store, cls, pop)
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, omega, lattice_q=q, arith=nb.ARITH_REFERENCE)
    done = it.runGraph(n, many=True) + it.runGraph(n, many=True) + it.runGraph(n, many=True)
    bk.syncAll()
    assert done == 3 * n + 2
    ref3 = oracle.run(q, pop, cls, mask, omega, done)
    assert np.array_equal(it.getInput().updateHostData().view(np.uint8), ref3.view(np.uint8))


def test_launch_chain_runs_out_of_counters_gracefully(nb, bk, oracle):
    """The launch chain takes its per-plane counters from a small pool (one slice per captured chain, 48 per device): the 49th
    and later captures find none and are recorded as ordinary step launches — same bits."""
    from neon_b200 import problems as P
    dim, q, iters = (40, 24, 20), 19, 4
    cls = oracle.classify(1, *dim)
    mask = oracle.wall_mask(q, cls)
    pop = oracle.init_pop(q, cls, np.float32)
    omega = oracle.omega_cavity(max(dim))
    grid = nb.dGrid(bk, dim)
    pop0, pop1, flag = P.setup_host(grid, q, np.float32, cls, pop)
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, omega, arith=nb.ARITH_REFERENCE)
    done = 0
    for k in range(56):
        it.__dict__.pop("_graphs", None)  # a fresh capture every time (the graphs, and their slices, stay alive in `keep`)
        done += it.runGraph(iters, many=True)
    bk.syncAll()
    ref = oracle.run(q, pop, cls, mask, omega, done)
    assert np.array_equal(it.getInput().updateHostData().view(np.uint8), ref.view(np.uint8))


def test_device_setup_matches_oracle(nb, bk, oracle):
    """nlbm_dense_classify / wall_mask / init_pop against the oracle, bit for bit, all geometries."""
    from neon_b200 import problems as P
    for geom in (0, 1, 2):
        for q, dt in ((19, np.float32), (27, np.float64), (19, np.float64), (27, np.float32)):
            nx, ny, nz = 37, 21, 18
            grid = nb.dGrid(bk, (nx, ny, nz))
            pop0, pop1, flag = P.setup_device(grid, q, dt, geom)
            bk.syncAll()
            cls = oracle.classify(geom, nx, ny, nz)
            assert np.array_equal(flag.classes(), cls)
            assert np.array_equal(flag.masks(), oracle.wall_mask(q, cls))
            ref = oracle.init_pop(q, cls, dt)
            assert np.array_equal(pop0.updateHostData().view(np.uint8), ref.view(np.uint8))
            assert np.array_equal(pop1.updateHostData().view(np.uint8), ref.view(np.uint8))


def test_non_bulk_cells_never_written(nb, bk, oracle):
    """LbmTools.h:304: only bulk cells are stored; wall cells keep their initial populations, padding stays zero."""
    from neon_b200 import problems as P
    grid = nb.dGrid(bk, (45, 20, 16))
    pop0, pop1, flag = P.setup_device(grid, 19, np.float32, 1)
    before = pop1.data.clone()
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, 1.2)
    it.run()
    bk.syncAll()
    cls = flag.classes()
    after = pop1.updateHostData()
    b4 = before.view(19, grid.nzm, 20, pop1.pitch_y)[:, :, :, :45].cpu().numpy()
    nonbulk = np.broadcast_to(cls != nb.BULK, after.shape)
    assert np.array_equal(after[nonbulk], b4[nonbulk])
    assert (after[~nonbulk] != b4[~nonbulk]).any()
    pad = pop1.view4[:, :, :, 45:].cpu().numpy()
    assert not pad.any()


def test_views_tile_the_partition(nb, bk, oracle):
    """INTERNAL + BOUNDARY == STANDARD, bit for bit, on a slab whose boundary planes hold bulk cells (BOUNDARY must
    cover z = 0 and z = nz-1)."""
    from neon_b200 import problems as P
    for part, kern in (((0, 3), nb.KERNEL_TMA), ((1, 3), nb.KERNEL_TMA), ((2, 3), nb.KERNEL_DIRECT), ((1, 3), nb.KERNEL_DIRECT)):
        grid = nb.dGrid(bk, (40, 24, 21), partition=part)
        pop0, pop1, flag = P.setup_device(grid, 19, np.float32, 1)
        pop0.data.uniform_(0.01, 0.1)  # ghost planes included: any data will do for this identity
        pop2 = grid.newField("pop2", 19, np.float32)
        pop2.data.copy_(pop1.data)
        a = nb.LbmContainers.iteration(nb.StencilSemantic.streaming, pop0, pop1, flag, 1.3, opts=nb.opt_kernel(kern))
        b = nb.LbmContainers.iteration(nb.StencilSemantic.streaming, pop0, pop2, flag, 1.3, opts=nb.opt_kernel(kern))
        a.run(0, nb.DataView.STANDARD)
        b.run(0, nb.DataView.INTERNAL)
        bk.syncAll()
        assert not torch.equal(pop1.data, pop2.data)
        b.run(0, nb.DataView.BOUNDARY)
        bk.syncAll()
        assert torch.equal(pop1.data, pop2.data)


@pytest.mark.parametrize("q,store,nparts", [(19, np.float32, 2), (19, np.float32, 3), (27, np.float64, 2), (19, np.float64, 4)])
def test_partitions_on_one_gpu_match_single_partition(nb, bk, oracle, q, store, nparts):
    """z-slab partitions + halo update reproduce the single-partition result bit for bit.  All
    partitions live on this one GPU, the way the reference tests multi-device (device list {0,0,0}); the face exchange
    is nlbm_dense_halo_push with the lattice semantic (5 of 19 / 9 of 27 populations), issued in OCC order:
    INTERNAL, halo, BOUNDARY."""
    from neon_b200 import _capi as capi
    from neon_b200 import problems as P
    dim, iters, omega = (36, 20, 23), 9, 1.25
    nx, ny, nz = dim
    cls = oracle.classify(1, nx, ny, nz)
    mask = oracle.wall_mask(q, cls)
    ref = oracle.run(q, oracle.init_pop(q, cls, store), cls, mask, omega, iters)
    parts = []
    for i in range(nparts):
        grid = nb.dGrid(bk, dim, partition=(i, nparts))
        pop0, pop1, flag = P.setup_device(grid, q, store, 1)
        assert np.array_equal(flag.masks(), mask[grid.z_origin:grid.z_origin + grid.nz_local])
        parts.append((grid, [pop0, pop1], flag))
    eb = np.dtype(store).itemsize
    st = bk.streamHandle(0)
    for t in range(iters):
        a, b = t & 1, (t & 1) ^ 1
        conts = [nb.LbmContainers.iteration(nb.StencilSemantic.streaming, p[a], p[b], f, omega, q, arith=nb.ARITH_REFERENCE)
                 for _, p, f in parts]
        for c in conts:
            c.run(0, nb.DataView.INTERNAL)
        for i in range(nparts - 1):  # faces between partition i (below) and i+1 (above)
            (g0, p0, _), (g1, p1, _) = parts[i], parts[i + 1]
            capi.call("nlbm_dense_halo_push", C.byref(g0.desc(p0[a], None, None)), p0[a].data.data_ptr(),
                      C.byref(g1.desc(p1[a], None, None)), p1[a].data.data_ptr(), eb, q, q, +1, st)
            capi.call("nlbm_dense_halo_push", C.byref(g1.desc(p1[a], None, None)), p1[a].data.data_ptr(),
                      C.byref(g0.desc(p0[a], None, None)), p0[a].data.data_ptr(), eb, q, q, -1, st)
        for c in conts:
            c.run(0, nb.DataView.BOUNDARY)
    bk.syncAll()
    got = np.concatenate([p[iters & 1].updateHostData() for _, p, _ in parts], axis=1)
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))


def test_rho_u_matches_oracle(nb, bk, oracle):
    from neon_b200 import problems as P
    for dt in (np.float32, np.float64):
        nx, ny, nz = 36, 20, 16
        cls = oracle.classify(1, nx, ny, nz)
        mask = oracle.wall_mask(19, cls)
        pop = oracle.run(19, oracle.init_pop(19, cls, dt), cls, mask, 1.4, 5)
        grid = nb.dGrid(bk, (nx, ny, nz))
        pop0, pop1, flag = P.setup_host(grid, 19, dt, cls, pop)
        rho, u = grid.newField("rho", 1, dt), grid.newField("u", 3, dt)
        nb.LbmContainers.computeRhoAndU(pop0, flag, rho, u).run(0)
        bk.syncAll()
        r_ref, u_ref = oracle.rho_u(pop, cls, mask)
        assert np.array_equal(rho.updateHostData()[0].view(np.uint8), r_ref.view(np.uint8))
        assert np.array_equal(u.updateHostData().view(np.uint8), u_ref.view(np.uint8))


def test_open_geometry_is_reported(nb, bk):
    """A bulk cell on the box edge has neighbours outside the domain: the reference reads invalid data there
    (SURVEY.md §8a row a6); the C layer counts them and the host side raises."""
    grid = nb.dGrid(bk, (16, 8, 8))
    flag = grid.newFlagField()
    flag.setClasses(np.full((8, 8, 16), nb.BULK, np.int32))
    with pytest.raises(nb.NeonException):
        flag.computeWallNghMask(19)


def test_error_codes(nb, bk):
    from neon_b200 import _capi as capi
    grid = nb.dGrid(bk, (16, 8, 8))
    f = grid.newField("p", 19, np.float32)
    flag = grid.newFlagField()
    with pytest.raises(nb.NeonException):
        nb.LbmContainers.iteration(nb.StencilSemantic.streaming, f, f, flag, 1.0)
    d = grid.desc(f, f, flag)
    assert capi.lib().nlbm_d3q19_f32_dense_step(C.byref(d), 1.0, 0, 0, None) == capi.ERR_INVALID
    assert "alias" in capi.last_error()
    g = grid.newField("g", 19, np.float32)
    d = grid.desc(f, g, flag)
    assert capi.lib().nlbm_d3q19_f32_dense_step(C.byref(d), 1.0, 7, 0, None) == capi.ERR_INVALID
    d.pitch_y = 17
    assert capi.lib().nlbm_d3q19_f32_dense_step(C.byref(d), 1.0, 0, 0, None) == capi.ERR_INVALID


def test_halo_pack_unpack_roundtrip(nb, bk):
    """pack(dir) of one partition + unpack(dir) into another moves exactly the crossing populations' boundary plane
    into the ghost plane (single process, two descriptors: the reference tests multi-device the same way by
    oversubscribing one GPU, libNeonDomain/tests/domain-halos/src/runHelper.h:64-67)."""
    from neon_b200 import _capi as capi
    from neon_b200.lattice import crossing
    nx, ny, nzl = 40, 12, 5
    for q, dt, tdt in ((19, np.float32, torch.float32), (27, np.float64, torch.float64)):
        d = capi.DenseDesc()
        d.nx, d.ny, d.nz_local, d.z_halo, d.gnx, d.gny, d.gnz = nx, ny, nzl, 1, nx, ny, 2 * nzl
        pb = C.c_size_t()
        capi.call("nlbm_dense_layout", C.byref(d), q, dt().itemsize, C.byref(pb), None)
        n = pb.value // dt().itemsize
        lower = torch.rand(n, dtype=tdt, device=bk.device)
        upper = torch.rand(n, dtype=tdt, device=bk.device)
        lo4 = lower.view(q, nzl + 2, ny, d.pitch_y)
        up4 = upper.view(q, nzl + 2, ny, d.pitch_y)
        up_before = upper.clone()
        nbytes = C.c_size_t()
        capi.call("nlbm_dense_halo_pack", C.byref(d), lower.data_ptr(), dt().itemsize, q, q, +1, None, C.byref(nbytes), None)
        assert nbytes.value == len(crossing(q, +1)) * d.pitch_z * dt().itemsize
        buf = torch.empty(nbytes.value // dt().itemsize, dtype=tdt, device=bk.device)
        st = bk.streamHandle(0)
        capi.call("nlbm_dense_halo_pack", C.byref(d), lower.data_ptr(), dt().itemsize, q, q, +1, buf.data_ptr(), None, st)
        capi.call("nlbm_dense_halo_unpack", C.byref(d), upper.data_ptr(), dt().itemsize, q, q, +1, buf.data_ptr(), st)
        bk.syncAll()
        expect = up_before.view(q, nzl + 2, ny, d.pitch_y).clone()
        for k in crossing(q, +1):
            expect[k, 0] = lo4[k, nzl]  # top local plane (zm = nz_local) -> lower ghost (zm = 0)
        assert torch.equal(up4, expect)
        # the direct (peer-store) variant does the same in one launch
        upper.copy_(up_before)
        capi.call("nlbm_dense_halo_push", C.byref(d), lower.data_ptr(), C.byref(d), upper.data_ptr(), dt().itemsize, q, q, +1, st)
        bk.syncAll()
        assert torch.equal(up4, expect)
        # downward, grid semantic (all components)
        lo_before = lower.clone()
        capi.call("nlbm_dense_halo_push", C.byref(d), upper.data_ptr(), C.byref(d), lower.data_ptr(), dt().itemsize, q, 0, -1, st)
        bk.syncAll()
        expect = lo_before.view(q, nzl + 2, ny, d.pitch_y).clone()
        expect[:, nzl + 1] = up4[:, 1]
        assert torch.equal(lo4, expect)


def test_skeleton_cuda_graph_replay(nb, bk, oracle):
    """A Skeleton captured into a CUDA graph gives the same bits as eager issue."""
    from neon_b200 import problems as P
    outs = []
    for graph in (False, True):
        grid = nb.dGrid(bk, (40, 24, 20))
        pop0, pop1, flag = P.setup_device(grid, 19, np.float32, 1)
        it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, 1.1, graph=graph)
        for _ in range(7):
            it.run()
        bk.syncAll()
        outs.append(it.getInput().updateHostData())
    assert np.array_equal(outs[0].view(np.uint8), outs[1].view(np.uint8))


def test_large_grid_properties(nb, bk, oracle):
    """256^3 (67 M cell updates per iteration): one oracle iteration as the checker, plus size-independent properties —
    views tile, vector widths agree, walls untouched, density stays near 1."""
    from neon_b200 import problems as P
    n = 256
    grid = nb.dGrid(bk, (n, n, n))
    pop0, pop1, flag = P.setup_device(grid, 19, np.float32, 0)
    omega = nb.omega_from_re(n)
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, omega,
                         arith=nb.ARITH_REFERENCE)
    it.run()
    bk.syncAll()
    got = it.getInput().updateHostData()
    cls = oracle.classify(0, n, n, n)
    assert np.array_equal(flag.classes(), cls)
    mask = oracle.wall_mask(19, cls)
    assert np.array_equal(flag.masks(), mask)
    init = oracle.init_pop(19, cls, np.float32)
    ref = init.copy()
    oracle.step(19, init, ref, cls, mask, omega)
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))
    del init, ref, got
    for _ in range(49):
        it.run()
    bk.syncAll()
    f = it.getInput().updateHostData().astype(np.float64)
    rho = f.sum(axis=0)[cls == nb.BULK]
    assert abs(rho.mean() - 1.0) < 1e-3 and rho.min() > 0.9 and rho.max() < 1.1


@pytest.mark.parametrize("q,store,parts", [(19, np.float32, 1), (27, np.float64, 1), (19, np.float32, 3)])
def test_x_face_cache_keeps_each_fields_own_wall_values(nb, bk, oracle, q, store, parts):
    """The x-face cache (nlbm_dense_wall_cache_build) is an optimisation, not a semantic: with it, without it and with wall
    values that DIFFER between the two fields, non-bulk cells keep exactly what their own field held and bulk cells get
    the same bits."""
    from neon_b200 import problems as P
    results = []
    for use_cache in (True, False):
        fields = []
        for part in range(parts):
            grid = nb.dGrid(bk, (44, 18, 15), partition=(part, parts))
            pop0, pop1, flag = P.setup_device(grid, q, store, 1)
            # make the wall values of pop1 differ from pop0's, x faces included (non-bulk cells only)
            nonbulk = torch.from_numpy(np.ascontiguousarray(flag.classes() != nb.BULK)).to(bk.device)
            loc = pop1.view4[:, grid.z_halo:grid.z_halo + grid.nz_local, :, :44]
            loc[:, nonbulk] = loc[:, nonbulk] * 0.5 + 0.125
            if use_cache:
                pop1.commitWalls()
            else:
                pop0.invalidateWalls()
                pop1.invalidateWalls()
            fields.append((grid, pop0, pop1, flag, pop0.updateHostData().copy(), pop1.updateHostData().copy()))
        for grid, pop0, pop1, flag, _, _ in fields:
            assert (pop1.wallCachePtr() is not None) == use_cache
            a = nb.LbmContainers.iteration(nb.StencilSemantic.streaming, pop0, pop1, flag, 1.3, q)
            b = nb.LbmContainers.iteration(nb.StencilSemantic.streaming, pop1, pop0, flag, 1.3, q)
            a.run(0, nb.DataView.STANDARD)  # ghost planes are not refreshed: the same (stale) data in both runs
            b.run(0, nb.DataView.STANDARD)
        bk.syncAll()
        out = []
        for grid, pop0, pop1, flag, p0_before, p1_before in fields:
            nb_mask = np.broadcast_to(flag.classes() != nb.BULK, p0_before.shape)
            p0, p1 = pop0.updateHostData(), pop1.updateHostData()
            assert np.array_equal(p0[nb_mask], p0_before[nb_mask]) and np.array_equal(p1[nb_mask], p1_before[nb_mask])
            assert not np.array_equal(p0[nb_mask], p1[nb_mask])
            out.append((p0, p1))
        results.append(out)
    for (a0, a1), (b0, b1) in zip(*results):
        assert np.array_equal(a0.view(np.uint8), b0.view(np.uint8)) and np.array_equal(a1.view(np.uint8), b1.view(np.uint8))
