"""The Skeleton's graph: dependencies from tokens, the OCC transformations (Occ none / standard / extended / twoWayExtended),
halo insertion, stream mapping — checked structurally on one process and numerically over gloo (2 and 3 ranks, CPU), with
map and stencil containers written here in torch (host-logic runtime: the scheduling code is the one the GPU runs use).

Reference behaviour: libNeonSkeleton/src/skeleton/internal/multiGpuGraph.cpp:43-70 (parse), :120-301 (OCC), :304-352 (halo
updates), libNeonSet/src/set/container/Graph.cpp:690-838 (streams, events); the reference's own tests of it:
libNeonSkeleton/tests/unit/sUt_skeleton (axpy / laplace sequences under every Occ, compared with a sequential run)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import neon_b200 as nb
from neon_b200.containers import Access, Container, Pattern, Token


def planes(grid, view):
    zh, nz = grid.z_halo, grid.nz_local
    if view == nb.DataView.STANDARD:
        return list(range(zh, zh + nz))
    if view == nb.DataView.INTERNAL:
        return list(range(zh + 1, zh + nz - 1))
    return sorted({zh, zh + nz - 1})


def map_container(name, src, dst, fn):
    g = src.grid

    def launch(streamIdx, view):
        for zm in planes(g, view):
            dst.view4[:, zm] = fn(src.view4[:, zm])
    return Container(name, [Token(src, Access.READ, Pattern.MAP), Token(dst, Access.WRITE, Pattern.MAP)], launch)


def axpy_container(name, a, x, y):
    """y = a * x + y: reads AND writes y"""
    g = x.grid

    def launch(streamIdx, view):
        for zm in planes(g, view):
            y.view4[:, zm] = a * x.view4[:, zm] + y.view4[:, zm]
    return Container(name, [Token(x, Access.READ, Pattern.MAP), Token(y, Access.READ, Pattern.MAP), Token(y, Access.WRITE, Pattern.MAP)], launch)


def stencil_container(name, src, dst):
    """dst[z] = src[z-1] + 2 src[z] + src[z+1], zero outside the box (ghost planes at the ends of the box are never filled)"""
    g = src.grid

    def launch(streamIdx, view):
        v = src.view4
        for zm in planes(g, view):
            lo = v[:, zm - 1] if zm - 1 >= 0 else 0
            hi = v[:, zm + 1] if zm + 1 < g.nzm else 0
            dst.view4[:, zm] = lo + 2 * v[:, zm] + hi
    return Container(name, [Token(src, Access.READ, Pattern.STENCIL, nb.StencilSemantic.standard), Token(dst, Access.WRITE, Pattern.MAP)], launch)


def np_stencil(a):
    out = 2 * a.copy()
    out[:, 1:] += a[:, :-1]
    out[:, :-1] += a[:, 1:]
    return out


def build(grid, occ, which):
    a, b, c, d = (grid.newField(n, 2, np.float64) for n in "abcd")
    if which == "map-stencil-map":
        ops = [map_container("M1", a, b, lambda t: 2 * t + 1), stencil_container("S", b, c), map_container("M2", c, d, lambda t: t - 3)]
    elif which == "two-stencils-then-overwrite":
        # both stencils read b (one halo update serves both); the last map overwrites b: WAR against both stencils
        ops = [map_container("M1", a, b, lambda t: t * t), stencil_container("S1", b, c), stencil_container("S2", b, d),
               map_container("M3", d, b, lambda t: -t)]
    elif which == "axpy-chain":
        ops = [axpy_container("A1", 0.5, a, b), stencil_container("S", b, c), axpy_container("A2", 2.0, c, d), stencil_container("S2", d, a)]
    else:
        raise ValueError(which)
    sk = nb.Skeleton(grid.backend)
    sk.sequence(ops, which, nb.Options(occ, nb.TransferMode.get))
    return (a, b, c, d), sk


def reference(which, A, B, C, D, runs):
    for _ in range(runs):
        if which == "map-stencil-map":
            B = 2 * A + 1
            C = np_stencil(B)
            D = C - 3
        elif which == "two-stencils-then-overwrite":
            B = A * A
            C = np_stencil(B)
            D = np_stencil(B)
            B = -D
        else:
            B = 0.5 * A + B
            C = np_stencil(B)
            D = 2.0 * C + D
            A = np_stencil(D)
    return A, B, C, D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


CASES = [(w, o) for w in ("map-stencil-map", "two-stencils-then-overwrite", "axpy-chain") for o in ("none", "standard", "extended", "twoWayExtended")]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bk = nb.Backend(runtime=nb.Runtime.openmp)
        dim = (6, 4, 13)
        for which, occ_name in CASES:
            grid = nb.dGrid(bk, dim)
            fields, sk = build(grid, getattr(nb.Occ, occ_name), which)
            rng = np.random.default_rng(5)
            glob = [rng.random((2,) + dim[::-1]) for _ in range(4)]
            for f, gdata in zip(fields, glob):
                f.view4[:, grid.z_halo:grid.z_halo + grid.nz_local, :, :dim[0]] = torch.from_numpy(gdata[:, grid.z_origin:grid.z_origin + grid.nz_local])
            runs = 3
            for _ in range(runs):
                sk.run()
            out = [f.gather() for f in fields]
            if rank == 0:
                ref = reference(which, *glob, runs)
                results[(which, occ_name)] = ([bool(np.allclose(o, r, rtol=0, atol=1e-12)) for o, r in zip(out, ref)], sk.schedule())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sequences_match_a_sequential_run(world):
    """Three sequences x four Occ modes, three runs each (the fields feed back), against numpy on the global arrays."""
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert len(results) == len(CASES)
    for case in CASES:
        ok, sched = results[case]
        assert all(ok), (case, ok, sched)


class _FakeBackend:
    """Enough of a Backend to build graphs for a rank of a 3-rank job without starting one."""
    world, rank, runtime = 3, 1, nb.Runtime.openmp
    group = None
    device = torch.device("cpu")

    def __init__(self):
        self._streams = []

    def setAvailableStreamSet(self, n):
        while len(self._streams) < n:
            self._streams.append(None)

    def newEvent(self):
        return None


def _graph(occ, which):
    bk = _FakeBackend()
    grid = nb.dGrid(bk, (6, 4, 12))
    _, sk = build(grid, occ, which)
    return sk


def test_standard_occ_splits_only_the_stencil():
    sk = _graph(nb.Occ.standard, "map-stencil-map")
    views = [(k, n, v) for _, k, n, v in sk.schedule() if k == "compute"]
    assert sorted(views) == sorted([("compute", "M1", "STANDARD"), ("compute", "S", "BOUNDARY"), ("compute", "S", "INTERNAL"),
                                    ("compute", "M2", "STANDARD")])
    d = sk.dependencies()
    halo = next(k for k in d if k[0] == "halo")
    assert d[halo] == [("compute", "M1", "STANDARD")]                       # the faces of b exist once M1 is done
    assert d[("compute", "S", "INTERNAL")] == [("compute", "M1", "STANDARD")]  # no ghost data needed: not behind the halo update
    assert d[("compute", "S", "BOUNDARY")] == [halo]
    assert d[("compute", "M2", "STANDARD")] == [("compute", "S", "BOUNDARY"), ("compute", "S", "INTERNAL")]
    # the halo update and the BOUNDARY half share a high-priority stream, the INTERNAL half stays on the main stream
    st = {(k, n, v): s for s, k, n, v in sk.schedule()}
    assert st[("compute", "S", "INTERNAL")] == 0 and st[halo] == st[("compute", "S", "BOUNDARY")] != 0


def test_extended_occ_starts_the_halo_update_after_the_boundary_half_of_the_map():
    sk = _graph(nb.Occ.extended, "map-stencil-map")
    d = sk.dependencies()
    halo = next(k for k in d if k[0] == "halo")
    assert ("compute", "M1", "BOUNDARY") in d and ("compute", "M1", "INTERNAL") in d
    assert d[halo] == [("compute", "M1", "BOUNDARY")]          # the update overlaps M1's INTERNAL half
    assert d[("compute", "S", "INTERNAL")] == [("compute", "M1", "BOUNDARY"), ("compute", "M1", "INTERNAL")]
    assert ("compute", "M2", "STANDARD") in d                  # the map behind the stencil is not split by Occ::extended
    order = [(k, n, v) for _, k, n, v in sk.schedule()]
    assert order.index(("compute", "M1", "BOUNDARY")) < order.index(("compute", "M1", "INTERNAL"))  # multiGpuGraph.cpp:190


def test_two_way_extended_occ_splits_both_sides():
    sk = _graph(nb.Occ.twoWayExtended, "map-stencil-map")
    d = sk.dependencies()
    # the map behind the stencil reads c with a MAP pattern: each half depends on the matching half of the stencil only
    assert d[("compute", "M2", "INTERNAL")] == [("compute", "S", "INTERNAL")]
    assert d[("compute", "M2", "BOUNDARY")] == [("compute", "S", "BOUNDARY")]


def test_one_halo_update_serves_two_readers_and_war_is_ordered():
    sk = _graph(nb.Occ.standard, "two-stencils-then-overwrite")
    d = sk.dependencies()
    assert sum(1 for k in d if k[0] == "halo") == 1            # b is not written between S1 and S2
    # M3 overwrites b, which both stencils read (their INTERNAL halves read the boundary planes too)
    assert set(d[("compute", "M3", "STANDARD")]) >= {("compute", "S2", "BOUNDARY"), ("compute", "S2", "INTERNAL")}
    reach = set()
    todo = [("compute", "M3", "STANDARD")]
    while todo:
        for p in d[todo.pop()]:
            if p not in reach:
                reach.add(p)
                todo.append(p)
    assert {("compute", "S1", "BOUNDARY"), ("compute", "S1", "INTERNAL")} <= reach


def test_single_partition_has_no_halo_and_no_split():
    bk = nb.Backend(runtime=nb.Runtime.openmp)
    grid = nb.dGrid(bk, (6, 4, 12))
    _, sk = build(grid, nb.Occ.twoWayExtended, "axpy-chain")
    assert [(s, k, v) for s, k, _, v in sk.schedule()] == [(0, "compute", "STANDARD")] * 4
