"""Model check of the two multi-PROCESS orderings (one process per GPU, no GPU needed here):

* `ipc` transport (neon_b200/ipc.py, IpcHalo.run) under the Skeleton's schedule: per halo update k every rank enqueues
  push -> signal(k) per neighbour, then wait(k) per neighbour; Occ::standard puts that next to the INTERNAL kernel;
* the fused step + face push kernel (FusedIteration / nlbm_dense_step_push): kernel t stores the crossing populations of its
  boundary planes into the neighbours' ghost planes of the field it writes and publishes t+1 once a boundary plane is
  complete; before launching kernel t a rank waits for the neighbours' value t.

CUDA events do not cross processes, the only cross-rank edges are "signal(v) happens before the wait for v returns".  The
model builds the happens-before relation from stream order, fork/join and those edges and requires every pair of conflicting
accesses (same rank, field and plane, at least one write; ghost planes are written by the neighbour) to be ordered, for 2-4
ranks over several iterations.  Deliberately broken variants must be caught.
"""
import itertools

import pytest

LO, HI = "ghost_lo", "ghost_hi"


class Model:
    def __init__(self):
        self.ops = []       # (rank, stream, reads, writes, label)
        self.edges = []
        self.last = {}
        self.signals = {}   # (receiver rank, slot, value) -> op id
        self.pending = []   # (op id, key) waits to link

    def op(self, rank, stream, label, reads=(), writes=()):
        i = len(self.ops)
        self.ops.append((rank, stream, frozenset(reads), frozenset(writes), label))
        p = self.last.get((rank, stream))
        if p is not None:
            self.edges.append((p, i))
        self.last[(rank, stream)] = i
        return i

    def fork(self, rank):
        a = self.op(rank, 0, "fork")
        b = self.op(rank, 1, "fork-wait")
        self.edges.append((a, b))

    def join(self, rank):
        a = self.op(rank, 1, "join")
        b = self.op(rank, 0, "join-wait")
        self.edges.append((a, b))

    def signal(self, rank, stream, receiver, slot, value):
        self.signals[(receiver, slot, value)] = self.op(rank, stream, f"signal {value}")

    def wait(self, rank, stream, slot, value):
        self.pending.append((self.op(rank, stream, f"wait {value}"), (rank, slot, value)))

    def unordered_hazards(self):
        edges = list(self.edges)
        for i, key in self.pending:
            if key in self.signals:
                edges.append((self.signals[key], i))
        m = len(self.ops)
        succ, indeg = [[] for _ in range(m)], [0] * m
        for a, b in edges:
            succ[a].append(b)
            indeg[b] += 1
        order, stack = [], [i for i in range(m) if indeg[i] == 0]
        while stack:
            i = stack.pop()
            order.append(i)
            for j in succ[i]:
                indeg[j] -= 1
                if indeg[j] == 0:
                    stack.append(j)
        assert len(order) == m, "cycle: the protocol would deadlock"
        reach = [0] * m
        for i in reversed(order):
            r = 0
            for j in succ[i]:
                r |= (1 << j) | reach[j]
            reach[i] = r
        touched = {}
        for i, (_, _, reads, writes, _) in enumerate(self.ops):
            for loc in reads:
                touched.setdefault(loc, []).append((i, False))
            for loc in writes:
                touched.setdefault(loc, []).append((i, True))
        bad = []
        for loc, acc in touched.items():
            for (a, wa), (b, wb) in itertools.combinations(acc, 2):
                if (wa or wb) and not ((reach[a] >> b) & 1 or (reach[b] >> a) & 1):
                    bad.append((loc, self.ops[a][4], self.ops[a][0], self.ops[b][4], self.ops[b][0]))
        return bad


def planes_read(nz, out):
    r = set()
    for z in out:
        for zz in (z - 1, z, z + 1):
            r.add(LO if zz < 0 else HI if zz >= nz else zz)
    return r


# ------------------------------------------------------------------------------------------------ ipc transport + Skeleton
def ipc_run(world, nz, occ, iters=6, signal_after_push=True, wait_before_boundary=True):
    m = Model()
    for t in range(iters):
        fin, fout, k = t % 2, 1 - t % 2, t + 1
        for r in range(world):
            nbrs = [(x, "from_below" if x > r else "from_above") for x in (r + 1, r - 1) if 0 <= x < world]
            hs = 1 if occ == "standard" else 0
            if occ == "standard":
                m.fork(r)
                m.op(r, 0, "INTERNAL", {(r, fin, p) for p in planes_read(nz, range(1, nz - 1))}, {(r, fout, z) for z in range(1, nz - 1)})
            for x, slot in nbrs:  # push my boundary plane into x's ghost plane, then publish k in x's flag word
                src, dst = (nz - 1, LO) if x > r else (0, HI)
                if not signal_after_push:
                    m.signal(r, hs, x, slot, k)
                m.op(r, hs, "push", {(r, fin, src)}, {(x, fin, dst)})
                if signal_after_push:
                    m.signal(r, hs, x, slot, k)
            if wait_before_boundary:
                for x, _ in nbrs:
                    m.wait(r, hs, "from_below" if x < r else "from_above", k)
            view = sorted({0, nz - 1}) if occ == "standard" else range(nz)
            m.op(r, hs, "BOUNDARY" if occ == "standard" else "STANDARD", {(r, fin, p) for p in planes_read(nz, view)},
                 {(r, fout, z) for z in view})
            if occ == "standard":
                m.join(r)
    return m.unordered_hazards()


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("occ", ["none", "standard"])
@pytest.mark.parametrize("nz", [2, 3, 6])
def test_ipc_transport_orders_every_hazard(world, occ, nz):
    assert ipc_run(world, nz, occ) == []


def test_ipc_broken_variants_are_caught():
    assert ipc_run(3, 4, "standard", signal_after_push=False) != []     # flag published before the data
    assert ipc_run(3, 4, "none", wait_before_boundary=False) != []      # nobody waits for the neighbours' faces


# ------------------------------------------------------------------------------------------------ fused step + face push
def fused_run(world, nz, iters=6, wait_before_launch=True, signal_after_plane=True):
    """kernel t of rank r, as three parts in the kernel's plane order (boundary planes first): plane 0 (reads ghost_lo, pushes
    down, publishes t+1 to the rank below), plane nz-1 (same upwards), interior.  The parts are concurrent: no edges between
    them inside one launch except that all follow the launch's waits and precede the next launch of the rank."""
    m = Model()
    for t in range(iters):
        fin, fout = t % 2, 1 - t % 2
        for r in range(world):
            dn, up = (r - 1 if r > 0 else None), (r + 1 if r < world - 1 else None)
            if t > 0 and wait_before_launch:
                for x, slot in ((dn, "from_below"), (up, "from_above")):
                    if x is not None:
                        m.wait(r, 0, slot, t)
            start = m.op(r, 0, f"launch {t}")
            parts = []
            for name, out, nbr, ghost_dst, slot in (("plane0", [0], dn, HI, "from_above"), ("planeTop", [nz - 1], up, LO, "from_below")):
                writes = {(r, fout, z) for z in out}
                if nbr is not None:
                    writes.add((nbr, fout, ghost_dst))
                i = len(m.ops)
                m.ops.append((r, 2, frozenset({(r, fin, p) for p in planes_read(nz, out)}), frozenset(writes), f"{name} {t}"))
                m.edges.append((start, i))
                parts.append(i)
                if nbr is not None:
                    s = len(m.ops)
                    m.ops.append((r, 2, frozenset(), frozenset(), f"signal {t + 1}"))
                    m.edges.append((i if signal_after_plane else start, s))  # broken variant: published before the plane is done
                    m.signals[(nbr, slot, t + 1)] = s
                    parts.append(s)
            if nz > 2:
                i = len(m.ops)
                inner = range(1, nz - 1)
                m.ops.append((r, 2, frozenset({(r, fin, p) for p in planes_read(nz, inner)}), frozenset({(r, fout, z) for z in inner}), f"interior {t}"))
                m.edges.append((start, i))
                parts.append(i)
            end = m.op(r, 0, f"end {t}")
            for i in parts:
                m.edges.append((i, end))
    return m.unordered_hazards()


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("nz", [2, 3, 6])
def test_fused_step_push_orders_every_hazard(world, nz):
    assert fused_run(world, nz) == []


def test_fused_broken_variants_are_caught():
    assert fused_run(3, 4, wait_before_launch=False) != []
    assert fused_run(3, 4, signal_after_plane=False) != []   # the counter must be published by the LAST warp of the plane


# ------------------------------------------------------------------------------------------------ pipelined halo (round 2)
def pipelined_run(world, nz, occ, iters=7, signal_after_push=True, wait_before_boundary=True, push_after_boundary=True):
    """Skeleton with Options.pipelinedHalo (neon_b200/skeleton.py): iteration t waits for the faces of ITS INPUT field (pushed by
    the neighbours right after their BOUNDARY kernel of iteration t-1; at t = 0 every rank pushes them first), runs BOUNDARY
    (high-priority side stream, issued before INTERNAL), then pushes the faces of its OUTPUT field and publishes the field's
    update counter in the neighbours' flag words.  Each field has its own flag words and counters."""
    m = Model()
    pushed = {}   # (rank, field) -> updates pushed
    waited = {}

    def push(r, s, f):
        pushed[(r, f)] = k = pushed.get((r, f), 0) + 1
        for x in (r + 1, r - 1):
            if not 0 <= x < world:
                continue
            src, dst = (nz - 1, LO) if x > r else (0, HI)
            slot = (f, "from_below" if x > r else "from_above")
            if not signal_after_push:
                m.signal(r, s, x, slot, k)
            m.op(r, s, f"push f{f} #{k}", {(r, f, src)}, {(x, f, dst)})
            if signal_after_push:
                m.signal(r, s, x, slot, k)

    def wait(r, s, f):
        waited[(r, f)] = k = waited.get((r, f), 0) + 1
        for x in (r + 1, r - 1):
            if 0 <= x < world:
                m.wait(r, s, (f, "from_below" if x < r else "from_above"), k)

    for t in range(iters):
        fin, fout = t % 2, 1 - t % 2
        for r in range(world):
            hs = 1 if occ == "standard" else 0
            if occ == "standard":
                m.fork(r)
            if pushed.get((r, fin), 0) == waited.get((r, fin), 0):
                push(r, hs, fin)
            if wait_before_boundary:
                wait(r, hs, fin)
            view = sorted({0, nz - 1}) if occ == "standard" else range(nz)
            if not push_after_boundary:
                push(r, hs, fout)
            m.op(r, hs, "BOUNDARY" if occ == "standard" else "STANDARD", {(r, fin, p) for p in planes_read(nz, view)},
                 {(r, fout, z) for z in view})
            if push_after_boundary:
                push(r, hs, fout)
            if occ == "standard":
                m.op(r, 0, "INTERNAL", {(r, fin, p) for p in planes_read(nz, range(1, nz - 1))}, {(r, fout, z) for z in range(1, nz - 1)})
                m.join(r)
    return m.unordered_hazards()


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("occ", ["none", "standard"])
@pytest.mark.parametrize("nz", [2, 3, 6])
def test_pipelined_halo_orders_every_hazard(world, occ, nz):
    assert pipelined_run(world, nz, occ) == []


def test_pipelined_broken_variants_are_caught():
    assert pipelined_run(3, 4, "standard", signal_after_push=False) != []    # flag published before the data
    assert pipelined_run(3, 4, "none", wait_before_boundary=False) != []     # nobody waits for the neighbours' faces
    assert pipelined_run(3, 4, "standard", push_after_boundary=False) != []  # faces pushed before they were computed
