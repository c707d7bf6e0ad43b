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


@dataclass
class Node:
    kind: str  # "fork" | "join" | "halo" | "compute"
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

    def sequence(self, operations: List[Container], name: str = "", options: Options = Options(), graph: bool = False) -> None:
        bk = self.backend
        self.name, self.options = name, options
        self.nodes = []
        multi = bk.world > 1
        for c in operations:
            halos = []
            if multi:
                for t in c.stencilReads():
                    halos.append(t.field.newHaloUpdate(t.semantic, options.transferMode, t.lattice_q,
                                                       getattr(c, "halo_transport", "auto")))
            if multi and halos and options.occ != Occ.none:
                self.nodes.append(Node("fork", "fork", 0))
                self.nodes.append(Node("compute", c.name, 0, DataView.INTERNAL, c))
                for h in halos:
                    self.nodes.append(Node("halo", h.name, 1, DataView.STANDARD, h))
                self.nodes.append(Node("compute", c.name, 1, DataView.BOUNDARY, c))
                self.nodes.append(Node("join", "join", 0))
            else:
                for h in halos:
                    self.nodes.append(Node("halo", h.name, 0, DataView.STANDARD, h))
                self.nodes.append(Node("compute", c.name, 0, DataView.STANDARD, c))
        bk.setAvailableStreamSet(1 + max((n.stream for n in self.nodes), default=0))
        self._use_graph = bool(graph) and not multi and bk.runtime == Runtime.stream
        self._graph = None

    def halos(self):
        """The halo-update containers this sequence inserted."""
        return [n.container for n in self.nodes if n.kind == "halo"]

    def schedule(self):
        """[(stream, kind, name, view)] in host issue order — what DB_multiGpuGraph.dot shows in the reference."""
        return [(n.stream, n.kind, n.name, n.view.name if n.view else None) for n in self.nodes]

    def _issue(self) -> None:
        bk = self.backend
        cuda = bk.runtime == Runtime.stream
        for n in self.nodes:
            if n.kind == "fork":
                if cuda:
                    e = bk.newEvent()
                    e.record(bk.stream(0))
                    bk.stream(1).wait_event(e)
            elif n.kind == "join":
                if cuda:
                    e = bk.newEvent()
                    e.record(bk.stream(1))
                    bk.stream(0).wait_event(e)
            else:
                n.container.run(n.stream, n.view)

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
