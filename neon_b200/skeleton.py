"""Skeleton — turns a sequence of Containers into a dependency graph, inserts halo updates, overlaps communication with
computation (OCC) and issues the result on the Backend's streams.

Mirrors libNeonSkeleton: Skeleton::sequence/run (include/Neon/skeleton/Skeleton.h:32-65), Options(Occ, TransferMode),
Occ (Occ.h:8-14) and MultiXpuGraph (src/skeleton/internal/multiGpuGraph.cpp):
  parse              tokens of every container -> RAW / WAR / WAW dependencies (:43-70, DependencyAnalyser)
  optimizations      Occ::standard        every stencil node is split into INTERNAL + BOUNDARY clones (:120-143)
                     Occ::extended        ... and so are the map nodes right in front of it, if ALL its predecessors are
                                          map nodes (:145-199): their BOUNDARY halves run first, so the halo update starts
                                          while their INTERNAL halves still compute
                     Occ::twoWayExtended  ... and the map nodes right behind it, if predecessors and successors
                                          qualify (:201-301)
  communications     a halo update in front of every stencil read whose consumer is not an INTERNAL clone (:304-352)
  scheduling         BFS levels, greedy stream mapping (a node takes the stream of a predecessor when it is free in its
                     level, else the first free one), one event per cross-stream edge, host issue order
                     (libNeonSet/src/set/container/Graph.cpp:652-661, 690-838, 992-1030)

Differences by design (B200-first):
  * dependencies between the halves of split nodes are derived from the CELLS each half touches (a MAP access of the
    INTERNAL half never meets the BOUNDARY half of its producer; a STENCIL access meets both) instead of cloning every
    edge, so the halo update of Occ::extended depends on the BOUNDARY half of the producer only;
  * ordering is by CUDA events only — the reference's halo update blocks the host on every device (SynchronizationContainer);
  * every stream but the first is a high-priority stream, nodes that are not INTERNAL / STANDARD compute prefer those, and
    they are issued first: the small BOUNDARY kernels and the face traffic never queue behind an INTERNAL kernel;
  * Options.pipelinedHalo: a field's faces are pushed right after the BOUNDARY half that wrote them, the consumer only waits;
  * the BOUNDARY view covers z_local in {0, nz-1} (the reference folds it onto {0,1});
  * with one device a whole run() can be captured once into a CUDA graph and replayed (launch-bound small domains).
Reductions (dot / norm containers) are outside the LBM path and are not modelled.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Dict, List, Optional, Set, Tuple

import torch

from .backend import Backend, Runtime
from .containers import Access, Container, Pattern
from .dgrid import DataView, TransferMode


class Occ(Enum):
    """Neon::skeleton::Occ (Occ.h:8-14)"""
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
    kind: str  # "halo" | "halo_wait" | "halo_push" | "compute"   (schedule() adds "fork" / "join" pseudo entries)
    name: str
    stream: int = 0
    view: Optional[DataView] = None
    container: Optional[Container] = None
    uid: int = -1
    op: int = -1                      # index of the container in the sequence (halo nodes: the consumer / producer)
    preds: Set[int] = field(default_factory=set)   # uids this node must run after
    level: int = 0


def _overlap(a: DataView, b: DataView) -> bool:
    """Do two views of the same partition share cells?"""
    return a == DataView.STANDARD or b == DataView.STANDARD or a == b


class Skeleton:
    def __init__(self, backend: Backend):
        self.backend = backend
        self.nodes: List[Node] = []   # in host issue order
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

    # ------------------------------------------------------------------------------------------------ graph construction
    @staticmethod
    def _pattern(c: Container) -> str:
        return "stencil" if c.stencilReads() else "map"

    @staticmethod
    def _op_dependencies(ops: List[Container]) -> List[Dict[int, List[Tuple[object, str, Pattern, Pattern]]]]:
        """deps[i][j] = [(field, "RAW"|"WAR"|"WAW", pattern of the earlier access, pattern of the later access)] for j < i:
        the data-dependency state machine of DependencyAnalyser (one record per field uid: last writer, readers since)."""
        deps: List[Dict[int, list]] = [dict() for _ in ops]
        last_write: Dict[int, Tuple[int, Pattern]] = {}
        readers: Dict[int, List[Tuple[int, Pattern]]] = {}
        for i, c in enumerate(ops):
            # reads first: a container that reads and writes the same field depends on the previous writer, not on itself
            for t in c.tokens:
                if t.access != Access.READ:
                    continue
                k = id(t.field)
                if k in last_write and last_write[k][0] != i:
                    deps[i].setdefault(last_write[k][0], []).append((t.field, "RAW", last_write[k][1], t.pattern))
            for t in c.tokens:
                if t.access != Access.WRITE:
                    continue
                k = id(t.field)
                for j, pat in readers.get(k, []):
                    if j != i:
                        deps[i].setdefault(j, []).append((t.field, "WAR", pat, t.pattern))
                if k in last_write and last_write[k][0] != i:
                    deps[i].setdefault(last_write[k][0], []).append((t.field, "WAW", last_write[k][1], t.pattern))
            for t in c.tokens:
                k = id(t.field)
                if t.access == Access.READ:
                    readers.setdefault(k, []).append((i, t.pattern))
            for t in c.tokens:
                k = id(t.field)
                if t.access == Access.WRITE:
                    last_write[k] = (i, t.pattern)
                    readers[k] = [(j, p) for j, p in readers.get(k, []) if j == i]
        return deps

    @staticmethod
    def _direct(deps: List[Dict[int, list]]):
        """Predecessors / successors after transitive reduction (Graph::removeRedundantDependencies)."""
        n = len(deps)
        reach: List[Set[int]] = [set() for _ in range(n)]
        direct: List[Set[int]] = [set() for _ in range(n)]
        for i in range(n):
            for j in sorted(deps[i], reverse=True):
                if j not in reach[i]:
                    direct[i].add(j)
                reach[i] |= {j} | reach[j]
        succ: List[Set[int]] = [set() for _ in range(n)]
        for i in range(n):
            for j in direct[i]:
                succ[j].add(i)
        return direct, succ

    def sequence(self, operations: List[Container], name: str = "", options: Options = Options(), graph: bool = False) -> None:
        bk = self.backend
        self.name, self.options = name, options
        multi = bk.world > 1
        ops = list(operations)
        pat = [self._pattern(c) for c in ops]
        deps = self._op_dependencies(ops)
        direct, succ = self._direct(deps)

        # ---- OCC: which containers are split into INTERNAL + BOUNDARY halves (multiGpuGraph.cpp:120-301)
        split: Set[int] = set()
        if multi and options.occ != Occ.none:
            for s in range(len(ops)):
                if pat[s] != "stencil" or ops[s].kind != "compute":
                    continue
                split.add(s)
                if options.occ == Occ.standard:
                    continue
                before = [j for j in direct[s] if ops[j].kind == "compute"]
                after = [j for j in succ[s] if ops[j].kind == "compute"]
                before_ok = bool(before) and len(before) == len(direct[s]) and all(pat[j] == "map" for j in before)
                after_ok = bool(after) and len(after) == len(succ[s]) and all(pat[j] == "map" for j in after)
                if options.occ == Occ.extended and before_ok:
                    split.update(before)
                if options.occ == Occ.twoWayExtended and before_ok and after_ok:
                    split.update(before)
                    split.update(after)

        nodes: List[Node] = []

        def add(kind, name_, view, container, op) -> Node:
            n = Node(kind, name_, 0, view, container, len(nodes), op)
            nodes.append(n)
            return n

        pieces: List[List[Node]] = []
        halo_in: List[List[Node]] = [[] for _ in ops]     # halo / halo_wait nodes in front of container i
        push_out: List[List[Node]] = [[] for _ in ops]    # halo_push nodes behind container i
        fresh: Dict[int, bool] = {}                        # field id -> its ghost planes are current within this sequence
        ghost_readers: Dict[int, List[Node]] = {}          # field id -> pieces that read the ghost planes since the last update
        for i, c in enumerate(ops):
            transport = getattr(c, "halo_transport", "auto")
            halos = []
            if multi and c.kind == "compute":
                for t in c.stencilReads():
                    if not fresh.get(id(t.field), False):
                        halos.append((t, self._halo_of(t.field, t.semantic, options.transferMode, t.lattice_q, transport)))
            pushes = []
            if multi and options.pipelinedHalo and halos and all(h.supportsSplit() for _, h in halos):
                for t in getattr(c, "push_after", []):
                    pushes.append((t, self._halo_of(t.field, t.semantic, options.transferMode, t.lattice_q, transport)))
                if not all(h.supportsSplit() for _, h in pushes):
                    pushes = []
            for t, h in halos:
                n = add("halo_wait" if pushes else "halo", h.name, DataView.STANDARD, h, i)
                halo_in[i].append(n)
                # the update overwrites ghost planes: after every earlier reader of them (WAR), after the producer's faces (RAW)
                n.preds |= {r.uid for r in ghost_readers.get(id(t.field), [])}
                for j in range(i - 1, -1, -1):
                    if any(w.access == Access.WRITE and w.field is t.field for w in ops[j].tokens):
                        n.preds |= {p.uid for p in pieces[j] if p.view != DataView.INTERNAL}
                        break
                fresh[id(t.field)] = True
                ghost_readers[id(t.field)] = []
            if i in split:
                pieces.append([add("compute", c.name, DataView.BOUNDARY, c, i), add("compute", c.name, DataView.INTERNAL, c, i)])
            else:
                pieces.append([add("compute" if c.kind == "compute" else c.kind, c.name, DataView.STANDARD, c, i)])
            # ---- dependencies of the halves on earlier halves, by the cells they touch
            for j, why in deps[i].items():
                for pi in pieces[i]:
                    for pj in pieces[j]:
                        for _, kind, early, late in why:
                            stencil = (late if kind == "RAW" else early) == Pattern.STENCIL
                            if stencil or _overlap(pi.view, pj.view):
                                pi.preds.add(pj.uid)
                                break
            for pi in pieces[i]:
                if pi.view != DataView.INTERNAL:
                    pi.preds |= {h.uid for h in halo_in[i]}
                    for t in c.stencilReads():
                        ghost_readers.setdefault(id(t.field), []).append(pi)
            for t in c.tokens:
                if t.access == Access.WRITE:
                    fresh[id(t.field)] = False
            for t, h in pushes:
                n = add("halo_push", h.name, DataView.STANDARD, h, i)
                n.preds |= {p.uid for p in pieces[i] if p.view != DataView.INTERNAL}
                push_out[i].append(n)

        self._schedule(nodes)
        bk.setAvailableStreamSet(1 + max((n.stream for n in self.nodes), default=0))
        self._use_graph = bool(graph) and not multi and bk.runtime == Runtime.stream
        self._graph = None
        cuda = bk.runtime == Runtime.stream
        # every event is created once: an iteration allocates nothing
        self._ev_fork = bk.newEvent() if cuda else None
        self._ev = {n.uid: bk.newEvent() for n in self.nodes if n.uid in self._signals} if cuda else {}

    # ------------------------------------------------------------------------------------------------------- scheduling
    def _schedule(self, nodes: List[Node]) -> None:
        """Levels, streams, events and issue order (Graph.cpp:652-661, 690-838)."""
        by_uid = {n.uid: n for n in nodes}
        # transitive reduction of the node graph
        reach: Dict[int, Set[int]] = {}
        for n in nodes:  # uids are in a topological order: every predecessor was created earlier
            kept, r = set(), set()
            for p in sorted(n.preds, reverse=True):
                if p not in r:
                    kept.add(p)
                r |= {p} | reach[p]
            n.preds, reach[n.uid] = kept, r
            n.level = 1 + max((by_uid[p].level for p in n.preds), default=-1)
        succs: Dict[int, Set[int]] = {n.uid: set() for n in nodes}
        for n in nodes:
            for p in n.preds:
                succs[p].add(n.uid)

        def main_lane(n: Node) -> bool:  # nodes that belong on the main (normal-priority) stream
            return n.kind == "compute" and n.view in (DataView.INTERNAL, DataView.STANDARD)

        # ---- streams, level by level (Graph.cpp:690-838): a node takes the stream of a predecessor if nobody in its level
        # holds it yet — a predecessor of its own lane first — else the first free stream of its lane.  Lanes: stream 0 is the
        # normal-priority main stream and carries INTERNAL / STANDARD compute nodes; halo nodes and BOUNDARY halves prefer
        # the high-priority streams 1.. whenever their level is shared with other nodes.
        for lvl in range(max((n.level for n in nodes), default=-1) + 1):
            todo = sorted((n for n in nodes if n.level == lvl), key=lambda n: (not main_lane(n), n.uid))
            booked: Set[int] = set()
            for n in todo:
                lane0 = main_lane(n) or len(todo) == 1
                mine = sorted((by_uid[p].stream for p in n.preds), key=lambda st: ((st == 0) != main_lane(n), st))
                s = next((st for st in mine if st not in booked), None)
                if s is None:
                    s = 0 if lane0 else 1
                    while s in booked:
                        s += 1
                booked.add(s)
                n.stream = s

        # ---- issue order: a topological order that prefers high-priority streams (their nodes are small and on the critical
        # path), then lower levels
        order, done = [], set()
        pending = list(nodes)
        while pending:
            ready = [n for n in pending if n.preds <= done]
            ready.sort(key=lambda n: (n.stream == 0, n.level, n.uid))
            n = ready[0]
            order.append(n)
            done.add(n.uid)
            pending.remove(n)
        self.nodes = order
        # ---- events: one per node with a successor on another stream; roots off the main stream wait for the fork event,
        # the last node of every side stream is joined into the main stream
        self._waits: Dict[int, List[int]] = {n.uid: [p for p in sorted(n.preds) if by_uid[p].stream != n.stream] for n in nodes}
        self._signals: Set[int] = {p for n in nodes for p in self._waits[n.uid]}
        self._fork_roots: Set[int] = {n.uid for n in nodes if n.stream != 0 and not any(by_uid[p].stream == n.stream for p in n.preds)
                                      and not self._waits[n.uid]}
        last_on: Dict[int, Node] = {}
        for n in order:
            last_on[n.stream] = n
        self._joins: List[int] = [n.uid for s, n in sorted(last_on.items()) if s != 0]
        self._signals |= set(self._joins)
        self._by_uid = by_uid

    def halos(self):
        """The halo-update containers this sequence uses."""
        out = []
        for n in self.nodes:
            if n.kind.startswith("halo") and n.container not in out:
                out.append(n.container)
        return out

    def schedule(self):
        """[(stream, kind, name, view)] in host issue order — what DB_multiGpuGraph.dot shows in the reference.  When more
        than one stream is used the list is bracketed by the fork / join of the side streams."""
        body = [(n.stream, n.kind, n.name, n.view.name if n.view else None) for n in self.nodes]
        if any(n.stream != 0 for n in self.nodes):
            return [(0, "fork", "fork", None)] + body + [(0, "join", "join", None)]
        return body

    def dependencies(self):
        """{(kind, name, view): sorted [(kind, name, view) of every direct predecessor]} — the scheduled graph."""
        key = lambda n: (n.kind, n.name, n.view.name if n.view else None)  # noqa: E731
        return {key(n): sorted(key(self._by_uid[p]) for p in n.preds) for n in self.nodes}

    # ------------------------------------------------------------------------------------------------------------ issue
    def _run_node(self, n: Node) -> None:
        if n.kind == "halo_wait":
            ipc = n.container._ipc_halo()
            if ipc.count == ipc.waited:  # nobody pushed the update this wait is for (first run): do it now
                n.container.push(n.stream)
            n.container.wait(n.stream)
        elif n.kind == "halo_push":
            n.container.push(n.stream)
        else:
            n.container.run(n.stream, n.view)

    def _issue(self, marks=None) -> None:
        bk = self.backend
        cuda = bk.runtime == Runtime.stream
        if cuda and self._fork_roots:
            self._ev_fork.record(bk.stream(0))
        for n in self.nodes:
            if cuda:
                if n.uid in self._fork_roots:
                    bk.stream(n.stream).wait_event(self._ev_fork)
                for p in self._waits[n.uid]:
                    bk.stream(n.stream).wait_event(self._ev[p])
            if marks is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(bk.stream(n.stream))
            self._run_node(n)
            if marks is not None:
                e1.record(bk.stream(n.stream))
                marks.append((n, e0, e1))
            if cuda and n.uid in self._signals:
                self._ev[n.uid].record(bk.stream(n.stream))
        if cuda:
            for uid in self._joins:
                bk.stream(0).wait_event(self._ev[uid])

    def timeline(self):
        """One run() with a CUDA event in front of and behind every node (on the node's stream): returns
        [(stream, kind, name, view, start_ms, end_ms)] relative to the start — the device-side picture nsys would draw of
        one iteration (which node overlapped which), measured without a profiler.  The extra events serialise nothing:
        they are recorded on the streams the nodes already use."""
        bk = self.backend
        assert bk.runtime == Runtime.stream
        base = torch.cuda.Event(enable_timing=True)
        base.record(bk.stream(0))
        for s in range(1, len(bk._streams)):
            bk.stream(s).wait_event(base)
        marks = []
        self._issue(marks)
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
