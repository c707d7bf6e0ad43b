"""Skeleton — schedules a sequence of Containers on the Backend's streams, inserting halo updates and overlapping
them with computation (OCC).

Mirrors libNeonSkeleton: Skeleton::sequence/run (include/Neon/skeleton/Skeleton.h:32-65), Options(Occ, TransferMode),
Occ (Occ.h:8-14) and MultiXpuGraph (src/skeleton/internal/multiGpuGraph.cpp): dependency analysis from tokens (:43-70),
OCC split of a stencil node into INTERNAL + BOUNDARY clones (:120-301), halo-update insertion on stencil-read edges
whose consumer is not INTERNAL (:304-352), stream mapping and event insertion (libNeonSet/src/set/container/Graph.cpp:
690-838) and sequential host issue (:992-1030).

Differences by design (B200-first):
  * ordering is by CUDA events only — the reference's halo update blocks the host on every device (SynchronizationContainer);
  * the BOUNDARY view covers z_local in {0, nz-1} (the reference folds it onto {0,1});
  * with one device a whole run() is captured once into a CUDA graph and replayed (launch-bound small domains).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import List, Optional

import torch

from .backend import Backend, Runtime
from .containers import Container
from .dgrid import DataView, TransferMode


class Occ(Enum):
    """Neon::skeleton::Occ (Occ.h:8-14).  For one stencil container per sequence — the LBM iteration — the extended
    variants schedule exactly like ``standard``."""
    none = "none"
    standard = "standard"
    extended = "extended"
    twoWayExtended = "twoWayExtended"


@dataclass
class Options:
    occ: Occ = Occ.none
    transferMode: TransferMode = TransferMode.get
    # Software-pipelined halo update (peer-store transport): the faces of a field are pushed into the neighbours' ghost
    # planes right AFTER the BOUNDARY kernel that wrote them (containers name those fields in ``push_after``), and the
    # consumer only waits for their arrival.  The transfer then overlaps the INTERNAL kernel of the SAME iteration and
    # the next iteration never waits for the wire (the reference pulls the faces right before the consumer,
    # multiGpuGraph.cpp:304-352, behind host-blocking syncs).
    pipelinedHalo: bool = False


@dataclass
class Node:
    kind: str  # "fork" | "join" | "halo" | "halo_wait" | "halo_push" | "compute"
    name: str
    stream: int
    view: Optional[DataView] = None
    container: Optional[Container] = None


class Skeleton:
    def __init__(self, backend: Backend):
        self.backend = backend
        self.nodes: List[Node] = []
        self.name = ""
        self._graph = None
        self._use_graph = False
        self._events = {}

    @staticmethod
    def _halo_of(field, semantic, transfer, lattice_q, transport):
        """One halo-update container per (field, semantic, ...): skeletons that exchange the same field share its flags."""
        cache = field.__dict__.setdefault("_halo_cache", {})
        key = (semantic, transfer, lattice_q, transport)
        if key not in cache:
            cache[key] = field.newHaloUpdate(semantic, transfer, lattice_q, transport)
        return cache[key]

    def sequence(self, operations: List[Container], name: str = "", options: Options = Options(), graph: bool = False) -> None:
        bk = self.backend
        self.name, self.options = name, options
        self.nodes = []
        multi = bk.world > 1
        SIDE = 1  # high-priority stream: halo + BOUNDARY
        for c in operations:
            halos = []
            transport = getattr(c, "halo_transport", "auto")
            if multi:
                for t in c.stencilReads():
                    halos.append(self._halo_of(t.field, t.semantic, options.transferMode, t.lattice_q, transport))
            pushes = []
            if multi and options.pipelinedHalo and halos and all(h.supportsSplit() for h in halos):
                for t in getattr(c, "push_after", []):
                    pushes.append(self._halo_of(t.field, t.semantic, options.transferMode, t.lattice_q, transport))
                if not all(h.supportsSplit() for h in pushes):
                    pushes = []
            pre = [Node("halo_wait", h.name, 0, DataView.STANDARD, h) for h in halos] if pushes else \
                  [Node("halo", h.name, 0, DataView.STANDARD, h) for h in halos]
            post = [Node("halo_push", h.name, 0, DataView.STANDARD, h) for h in pushes]
            if multi and halos and options.occ != Occ.none:
                # the side stream's nodes are issued FIRST and run at high priority: BOUNDARY and the face traffic are
                # out of the way while INTERNAL still fills the chip (round 1 issued INTERNAL first: the side stream's
                # work then ran in INTERNAL's tail and OCC bought nothing, profiles/r01l)
                self.nodes.append(Node("fork", "fork", 0))
                for n in pre + [Node("compute", c.name, SIDE, DataView.BOUNDARY, c)] + post:
                    n.stream = SIDE
                    self.nodes.append(n)
                self.nodes.append(Node("compute", c.name, 0, DataView.INTERNAL, c))
                self.nodes.append(Node("join", "join", 0))
            else:
                self.nodes += pre + [Node("compute", c.name, 0, DataView.STANDARD, c)] + post
        bk.setAvailableStreamSet(1 + max((n.stream for n in self.nodes), default=0))
        self._use_graph = bool(graph) and not multi and bk.runtime == Runtime.stream
        self._graph = None
        cuda = bk.runtime == Runtime.stream
        self._ev_fork = bk.newEvent() if cuda else None  # created once: an iteration allocates nothing
        self._ev_join = bk.newEvent() if cuda else None

    def halos(self):
        """The halo-update containers this sequence uses."""
        out = []
        for n in self.nodes:
            if n.kind.startswith("halo") and n.container not in out:
                out.append(n.container)
        return out

    def schedule(self):
        """[(stream, kind, name, view)] in host issue order — what DB_multiGpuGraph.dot shows in the reference."""
        return [(n.stream, n.kind, n.name, n.view.name if n.view else None) for n in self.nodes]

    def _issue(self) -> None:
        bk = self.backend
        cuda = bk.runtime == Runtime.stream
        for n in self.nodes:
            if n.kind == "fork":
                if cuda:
                    self._ev_fork.record(bk.stream(0))
                    bk.stream(1).wait_event(self._ev_fork)
            elif n.kind == "join":
                if cuda:
                    self._ev_join.record(bk.stream(1))
                    bk.stream(0).wait_event(self._ev_join)
            elif n.kind == "halo_wait":
                ipc = n.container._ipc_halo()
                if ipc.count == ipc.waited:  # nobody pushed the update this wait is for (first run): do it now
                    n.container.push(n.stream)
                n.container.wait(n.stream)
            elif n.kind == "halo_push":
                n.container.push(n.stream)
            else:
                n.container.run(n.stream, n.view)

    def timeline(self):
        """One run() with a CUDA event in front of and behind every node (on the node's stream): returns
        [(stream, kind, name, view, start_ms, end_ms)] relative to the first node's start — the device-side picture nsys
        would draw of one iteration (which node overlapped which), measured without a profiler.  The extra events
        serialise nothing: they are recorded on the streams the nodes already use."""
        bk = self.backend
        assert bk.runtime == Runtime.stream
        marks = []
        base = torch.cuda.Event(enable_timing=True)
        base.record(bk.stream(0))
        for s in range(1, len(bk._streams)):
            bk.stream(s).wait_event(base)
        saved = self.nodes
        for n in saved:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(bk.stream(n.stream))
            self.nodes = [n]
            self._issue()
            e1.record(bk.stream(n.stream))
            marks.append((n, e0, e1))
        self.nodes = saved
        bk.syncAll()
        return [(n.stream, n.kind, n.name, n.view.name if n.view else None, base.elapsed_time(e0), base.elapsed_time(e1))
                for n, e0, e1 in marks]

    def run(self) -> None:
        if not self._use_graph:
            self._issue()
            return
        bk = self.backend
        if self._graph is None:
            # capture on the main stream itself so that the C layer's launches (made on that stream) are recorded
            main = bk.stream(0)
            g = torch.cuda.CUDAGraph()
            main.synchronize()
            with torch.cuda.stream(main):
                g.capture_begin()
                try:
                    self._issue()
                finally:
                    g.capture_end()
            self._graph = g
        with torch.cuda.stream(bk.stream(0)):
            self._graph.replay()
