"""Model check of the event-only ordering of the C++ veneer's multi-GPU iteration (no GPU).

One process drives n devices; per device two streams; a halo update is `record ready` -> wait for the neighbours' `ready` ->
face transfers (get: pull into my ghost planes, put: push into the neighbours') -> (put) `record done` / wait for the
neighbours' `done` (neon_b200/cpp/include/Neon/domain/dGrid.h, DenseHaloImpl::run); Occ::standard runs INTERNAL on stream 0
next to {halo update -> BOUNDARY} on stream 1 between a fork and a join (Skeleton.h, Skeleton::issue).  The reference fences the
same transfers with host-blocking syncs (SynchronizationContainer.h:37-42); here nothing blocks the host, so every hazard of
SURVEY.md §8e must be ordered by the happens-before relation the streams and events create.  This test replays the host's issue
order for several iterations (fields alternate), builds that relation, and checks every pair of conflicting accesses
(same device, field and plane, at least one write) — including the cross-device ones — and shows that two deliberately
broken variants (no `ready` wait; put without the `done` wait) are caught.
"""
import itertools

import pytest

GHOST_LO, GHOST_HI = "ghost_lo", "ghost_hi"


class Model:
    def __init__(self, n_dev, nz, occ, mode, wait_ready=True, wait_done=True):
        self.n, self.nz, self.occ, self.mode = n_dev, nz, occ, mode
        self.wait_ready, self.wait_done = wait_ready, wait_done
        self.ops = []          # (id, device, stream, reads, writes, label); reads/writes: sets of (device, field, plane)
        self.edges = set()     # happens-before edges between op ids
        self.last_on_stream = {}
        self.last_record = {}  # event name -> op id of the most recent record (what a later cudaStreamWaitEvent sees)

    def op(self, dev, stream, label, reads=(), writes=()):
        i = len(self.ops)
        self.ops.append((i, dev, stream, frozenset(reads), frozenset(writes), label))
        prev = self.last_on_stream.get((dev, stream))
        if prev is not None:
            self.edges.add((prev, i))  # stream order
        self.last_on_stream[(dev, stream)] = i
        return i

    def record(self, dev, stream, event):
        self.last_record[event] = self.op(dev, stream, f"record {event}")

    def wait(self, dev, stream, event):
        i = self.op(dev, stream, f"wait {event}")
        if event in self.last_record:
            self.edges.add((self.last_record[event], i))

    # ---- what the kernels touch ------------------------------------------------------------------------------------
    def compute(self, dev, stream, view, fin, fout):
        nz = self.nz
        if view == "STANDARD":
            out = range(nz)
        elif view == "INTERNAL":
            out = range(1, nz - 1)
        else:
            out = sorted({0, nz - 1})
        reads = set()
        for z in out:
            for zz in (z - 1, z, z + 1):
                plane = GHOST_LO if zz < 0 else GHOST_HI if zz >= nz else zz
                reads.add((dev, fin, plane))
        self.op(dev, stream, f"{view} t", reads, {(dev, fout, z) for z in out})

    def halo(self, stream, field):
        n, nz = self.n, self.nz
        for d in range(n):
            self.record(d, stream, f"ready{d}")
        for d in range(n):
            nbrs = [x for x in (d - 1, d + 1) if 0 <= x < n]
            if self.wait_ready:
                for x in nbrs:
                    self.wait(d, stream, f"ready{x}")
            for x in nbrs:
                if self.mode == "get":   # pull the neighbour's boundary plane into my ghost plane
                    src_plane, dst_plane = (nz - 1, GHOST_LO) if x < d else (0, GHOST_HI)
                    self.op(d, stream, "pull", {(x, field, src_plane)}, {(d, field, dst_plane)})
                else:                    # push my boundary plane into the neighbour's ghost plane
                    src_plane, dst_plane = (nz - 1, GHOST_LO) if x > d else (0, GHOST_HI)
                    self.op(d, stream, "push", {(d, field, src_plane)}, {(x, field, dst_plane)})
        if self.mode == "put":
            for d in range(n):
                self.record(d, stream, f"done{d}")
            if self.wait_done:
                for d in range(n):
                    for x in (d - 1, d + 1):
                        if 0 <= x < n:
                            self.wait(d, stream, f"done{x}")

    def iteration(self, t):
        fin, fout = t % 2, 1 - t % 2
        n = self.n
        if self.occ == "none":
            self.halo(0, fin)
            for d in range(n):
                self.compute(d, 0, "STANDARD", fin, fout)
        else:
            for d in range(n):   # fork
                self.record(d, 0, f"fork{d}")
                self.wait(d, 1, f"fork{d}")
            for d in range(n):
                self.compute(d, 0, "INTERNAL", fin, fout)
            self.halo(1, fin)
            for d in range(n):
                self.compute(d, 1, "BOUNDARY", fin, fout)
            for d in range(n):   # join
                self.record(d, 1, f"join{d}")
                self.wait(d, 0, f"join{d}")

    # ---- happens-before closure and the hazard check ---------------------------------------------------------------
    def unordered_hazards(self):
        m = len(self.ops)
        succ = [[] for _ in range(m)]
        for a, b in self.edges:
            succ[a].append(b)
        reach = [0] * m           # bitset of ops reachable from i; edges go forward in issue order
        for i in range(m - 1, -1, -1):
            r = 0
            for j in succ[i]:
                r |= (1 << j) | reach[j]
            reach[i] = r
        touched = {}
        for i, _, _, reads, writes, _ in self.ops:
            for loc in reads:
                touched.setdefault(loc, []).append((i, False))
            for loc in writes:
                touched.setdefault(loc, []).append((i, True))
        bad = []
        for loc, acc in touched.items():
            for (a, wa), (b, wb) in itertools.combinations(acc, 2):
                if (wa or wb) and a != b and not (reach[a] >> b) & 1:
                    bad.append((loc, self.ops[a][5], self.ops[a][1], self.ops[b][5], self.ops[b][1]))
        return bad


def run(n_dev, nz, occ, mode, iters=5, **kw):
    m = Model(n_dev, nz, occ, mode, **kw)
    for t in range(iters):
        m.iteration(t)
    return m.unordered_hazards()


@pytest.mark.parametrize("n_dev", [2, 3, 4, 8])
@pytest.mark.parametrize("occ", ["none", "standard"])
@pytest.mark.parametrize("mode", ["get", "put"])
@pytest.mark.parametrize("nz", [2, 3, 5])
def test_every_hazard_is_ordered_by_events(n_dev, occ, mode, nz):
    assert run(n_dev, nz, occ, mode) == []


@pytest.mark.parametrize("occ", ["none", "standard"])
def test_dropping_the_ready_wait_is_caught(occ):
    """get: without waiting for the neighbour's `ready`, a pull may read a boundary plane the neighbour is still writing (RAW)
    and the neighbour may overwrite it while it is read (WAR).  put: the `done` waits of the previous update already order
    the push behind the neighbour's last reads of its ghost plane, so the `ready` wait is redundant there (kept: it is one
    cudaStreamWaitEvent and makes the two modes symmetric)."""
    assert run(3, 4, occ, "get", wait_ready=False) != []
    assert run(3, 4, occ, "put", wait_ready=False) == []


@pytest.mark.parametrize("occ", ["none", "standard"])
def test_put_without_the_done_wait_is_caught(occ):
    """in put mode the ghost planes are written by the NEIGHBOUR's stream: the consumer has to wait for its `done`"""
    bad = run(3, 4, occ, "put", wait_done=False)
    assert bad and any(a == "push" or b == "push" for _, a, _, b, _ in bad)
