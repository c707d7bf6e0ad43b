"""Host-side logic without a GPU: partitioning, lattice tables, set-up arrays, Skeleton schedules, and the N>1 halo
exchange over gloo with world_size 2 (fields in host memory; compute containers are not run)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import neon_b200 as nb
from neon_b200 import problems as P
from neon_b200.lattice import D3Q19, D3Q27, crossing


def test_partition_rule():
    """dGrid_imp.h:43-62: floor(Z/n) each, the first Z mod n partitions take one more plane."""
    assert nb.partition_z(10, 3) == ([4, 3, 3], [0, 4, 7])
    assert nb.partition_z(1024, 8) == ([128] * 8, [128 * i for i in range(8)])
    assert nb.partition_z(7, 7)[0] == [1] * 7
    with pytest.raises(ValueError):
        nb.partition_z(3, 4)


def test_lattices_match_the_oracle_tables(oracle):
    for L in (D3Q19, D3Q27):
        c, opp, w = oracle.tables(L.Q)
        assert np.array_equal(L.c, c) and np.array_equal(L.opp, opp) and np.array_equal(L.t, w)
    assert crossing(19, +1) == [6, 8, 12, 15, 17] and crossing(19, -1) == [2, 5, 7, 16, 18]
    assert len(crossing(27, +1)) == 9 and len(crossing(27, -1)) == 9


@pytest.mark.parametrize("geom", [0, 1, 2])
def test_host_setup_matches_oracle(oracle, geom):
    dim = (21, 13, 17)
    cls = P.host_classes(geom, dim)
    assert np.array_equal(cls, oracle.classify(geom, *dim))
    for q in (19, 27):
        for dt in (np.float32, np.float64):
            assert np.array_equal(P.host_populations(q, cls, dt).view(np.uint8), oracle.init_pop(q, cls, dt).view(np.uint8))


def test_single_partition_schedule():
    bk = nb.Backend(runtime=nb.Runtime.openmp)
    grid = nb.dGrid(bk, (16, 8, 8))
    a, b = grid.newField("a", 19, np.float32), grid.newField("b", 19, np.float32)
    flag = grid.newFlagField()
    it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.standard, nb.TransferMode.get, a, b, flag, 1.0)
    # one device: Begin -> LBM(STANDARD) -> End (multiGpuGraph.cpp:306 returns before inserting communications)
    assert it.lbmTwoPop[0].schedule() == [(0, "compute", "LBM_iteration_D3Q19", "STANDARD")]
    assert it.getInput() is a and it.getOutput() is b


def test_field_layout_and_roundtrip():
    bk = nb.Backend(runtime=nb.Runtime.openmp)
    grid = nb.dGrid(bk, (20, 6, 5))
    f = grid.newField("f", 3, np.float64)
    assert f.pitch_y == 64 and f.pitch_z == 64 * 6 and f.view4.shape == (3, 5, 6, 64)
    host = np.random.default_rng(0).random((3, 5, 6, 20))
    f.updateDeviceData(host)
    assert np.array_equal(f.updateHostData(), host)
    assert not f.view4[..., 20:].any()


def test_ranks_upload_their_own_part_of_the_host_mirror():
    """one process per GPU: a rank holds only the planes it needs (slab + in-box ghost planes) of the host arrays;
    updateDeviceData / setClasses with host_z0 fill the partition exactly like the global arrays do (bench.py e2e at N > 1)"""
    bk = nb.Backend(runtime=nb.Runtime.openmp)
    dim, q = (20, 6, 11), 5
    rng = np.random.default_rng(1)
    glob = rng.random((q,) + dim[::-1]).astype(np.float32)
    cls = rng.integers(0, 3, dim[::-1]).astype(np.int32)
    for part in range(3):
        g = nb.dGrid(bk, dim, partition=(part, 3))
        lo, hi = max(0, g.z_origin - g.z_halo), min(dim[2], g.z_origin + g.nz_local + g.z_halo)
        a, b = g.newField("a", q, np.float32), g.newField("b", q, np.float32)
        a.updateDeviceData(glob)
        b.updateDeviceData(np.ascontiguousarray(glob[:, lo:hi]), host_z0=lo)
        assert torch.equal(a.data, b.data)
        fa, fb = g.newFlagField(like=a), g.newFlagField(like=a)
        fa.setClasses(cls)
        fb.setClasses(np.ascontiguousarray(cls[lo:hi]), host_z0=lo)
        assert torch.equal(fa.words, fb.words)
        assert np.array_equal(fa.classes(), cls[g.z_origin:g.z_origin + g.nz_local])
        with pytest.raises(AssertionError):
            b.updateDeviceData(np.ascontiguousarray(glob[:, lo + 1:hi]), host_z0=lo + 1)  # does not cover the partition


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, semantic, q, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bk = nb.Backend(runtime=nb.Runtime.openmp)
        dim = (12, 5, 11)
        grid = nb.dGrid(bk, dim)
        assert grid.z_halo == 1 and grid.nz_local == nb.partition_z(11, world)[0][rank]
        f = grid.newField("pop", q, np.float32)
        # global field whose value encodes (component, z, y, x): every rank uploads its slab WITHOUT ghosts
        zz, yy, xx = np.meshgrid(np.arange(dim[2]), np.arange(dim[1]), np.arange(dim[0]), indexing="ij")
        glob = np.stack([(k * 1000 + zz * 100 + yy * 10 + xx).astype(np.float32) for k in range(q)])
        f.view4[:, 1:1 + grid.nz_local, :, :dim[0]] = torch.from_numpy(glob[:, grid.z_origin:grid.z_origin + grid.nz_local])
        sem = nb.StencilSemantic.streaming if semantic == "streaming" else nb.StencilSemantic.standard
        flag = grid.newFlagField()
        g = grid.newField("out", q, np.float32)
        it = nb.LbmIteration(sem, nb.Occ.standard, nb.TransferMode.get, f, g, flag, 1.0, lattice_q=q)
        sched = it.lbmTwoPop[0].schedule()
        # OCC: INTERNAL on stream 0, halo + BOUNDARY on stream 1, joined (multiGpuGraph.cpp:120-143, 304-352); the side
        # stream's nodes are issued first (they run at high priority while INTERNAL fills the chip)
        assert [(s, k, v) for s, k, _, v in sched] == [(0, "fork", None), (1, "halo", "STANDARD"), (1, "compute", "BOUNDARY"),
                                                       (0, "compute", "INTERNAL"), (0, "join", None)]
        none = nb.Skeleton(bk)
        none.sequence([nb.LbmContainers.iteration(sem, f, g, flag, 1.0, q)], "noOcc", nb.Options(nb.Occ.none, nb.TransferMode.get))
        assert [(s, k, v) for s, k, _, v in none.schedule()] == [(0, "halo", "STANDARD"), (0, "compute", "STANDARD")]
        halo = f.newHaloUpdate(sem, nb.TransferMode.get, q)
        halo.run(0)
        lo, hi = f.view4[:, 0, :, :dim[0]].numpy(), f.view4[:, grid.nz_local + 1, :, :dim[0]].numpy()
        up_set = range(q) if semantic == "standard" else crossing(q, +1)
        dn_set = range(q) if semantic == "standard" else crossing(q, -1)
        ok = True
        for k in range(q):
            if rank > 0:  # lower ghost = plane z_origin-1 of the components that move up
                want = glob[k, grid.z_origin - 1] if k in up_set else 0
                ok &= bool(np.array_equal(lo[k], np.broadcast_to(want, lo[k].shape)))
            else:
                ok &= not lo[k].any()
            if rank < world - 1:
                want = glob[k, grid.z_origin + grid.nz_local] if k in dn_set else 0
                ok &= bool(np.array_equal(hi[k], np.broadcast_to(want, hi[k].shape)))
            else:
                ok &= not hi[k].any()
        nbytes = halo.bytesPerDirection(+1)
        results[rank] = (ok, nbytes)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("semantic,q,world", [("streaming", 19, 2), ("standard", 19, 2), ("streaming", 27, 3)])
def test_halo_update_over_gloo(semantic, q, world):
    """SoA halo update across ranks (libNeonDomain/tests/domain-halos: global coordinates written into the field, halo
    update, z+-1 neighbours checked) — here per semantic: the lattice semantic moves 5 of 19 / 9 of 27 components."""
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), semantic, q, results), nprocs=world, join=True)
    assert all(results[r][0] for r in range(world)), dict(results)
    ncomp = q if semantic == "standard" else (5 if q == 19 else 9)
    assert results[0][1] == ncomp * 128 * 5 * 4
