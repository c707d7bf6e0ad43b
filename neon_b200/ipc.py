"""Peer-store halo transport between processes: CUDA-IPC mappings + device-side flags.

The reference drives every GPU from one process, so a halo update is a ``cudaMemcpyPeerAsync`` between two of its own
allocations (libNeonSet/src/set/DevSet.cpp:401-437) ordered by its own events and host-blocking syncs.  Here every GPU
belongs to its own process.  Once, at first use, each rank exports its field (and a few flag words) through CUDA IPC to
its z-neighbours; afterwards a halo update is, per neighbour,

    nlbm_dense_halo_push   my boundary plane (crossing populations only) -> the neighbour's ghost plane, over NVLink
    nlbm_flag_signal       publish my update counter in the neighbour's flag word (after the copy, stream order)
    nlbm_flag_wait         hold my stream until the neighbour's counter for this update arrived

all enqueued on the stream the Skeleton gave the halo node — no host synchronisation, no staging buffer, no NCCL call on
the data path.  Hazards: RAW — the ghost plane is complete before my BOUNDARY kernel because that
kernel follows the wait in stream order; WAR — a neighbour overwrites my ghost plane of field A for update k+1 only
after its own BOUNDARY kernel of the iteration in between, which waited for my signal of that iteration, which I enqueue
after the BOUNDARY kernel that read the plane.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _capi as capi

_FLAG_WORDS = 32
FROM_BELOW, FROM_ABOVE = 0, 16  # flag word written by the rank below / above (different 64-byte lines)
TIMEOUT_MS = 20000

_imported = {}  # handle bytes -> base address of the mapping in this process (a handle may be opened once per process)


def export_ptr(ptr: int):
    """(handle bytes, offset) of the device allocation that contains ``ptr`` — picklable."""
    h = C.create_string_buffer(64)
    off = C.c_uint64()
    capi.call("nlbm_ipc_export", C.c_void_p(ptr), h, C.byref(off))
    return bytes(h.raw), int(off.value)


def import_ptr(handle: bytes, offset: int) -> int:
    """Address, in this process, of what ``export_ptr`` described in another one (same node)."""
    if handle not in _imported:
        base = C.c_void_p()
        capi.call("nlbm_ipc_import", C.create_string_buffer(handle, 64), C.byref(base))
        _imported[handle] = int(base.value)
    return _imported[handle] + offset


def close_all() -> None:
    """Unmaps every neighbour allocation this process imported.  Collective in spirit: call it on every rank (between
    barriers) BEFORE the exporting ranks free the fields, and drop the iterations/halo updates that used the mappings."""
    for base in _imported.values():
        capi.call("nlbm_ipc_close", C.c_void_p(base))
    _imported.clear()


class IpcHalo:
    def __init__(self, halo):
        self.halo = halo
        f = halo.field
        g = f.grid
        bk = g.backend
        self.count = 0   # updates pushed
        self.waited = 0  # updates waited for
        self.counter = torch.zeros(1, dtype=torch.int32, device=bk.device)
        self.flags = torch.zeros(_FLAG_WORDS, dtype=torch.int32, device=bk.device)
        self.err = torch.zeros(1, dtype=torch.int32, device=bk.device)
        torch.cuda.synchronize(bk.device)
        # every rank publishes (field, flags) handles; each maps only its neighbours' allocations
        self.block = getattr(g, "kind", "dense") == "block"
        shape = (g.n_blocks, g.n_blocks_alloc, g.n_ghost_down, g.n_down, g.n_up) if self.block else None
        mine = (export_ptr(f.data.data_ptr()), export_ptr(self.flags.data_ptr()), shape)
        handles = [None] * bk.world
        dist.all_gather_object(handles, mine, group=bk.group)
        dn, up = g.neighbours()
        self.peer = {}
        for nbr in (dn, up):
            if nbr is None:
                continue
            hf, hg, _ = handles[nbr]
            self.peer[nbr] = (import_ptr(*hf), import_ptr(*hg))  # addresses aliasing the neighbour's memory
        self._descs = self._neighbour_block_descs(handles) if self.block else self._neighbour_descs()
        bk.barrier()

    def _neighbour_descs(self):
        """Descriptor of each neighbour's partition (same box, its own slab height)."""
        f = self.halo.field
        g = f.grid
        out = {}
        for nbr in self.peer:
            d = g.desc(f, None, None).clone()
            d.nz_local = g.sizes[nbr]
            d.z_origin = g.origins[nbr]
            d.pitch_q = d.pitch_z * (d.nz_local + 2 * d.z_halo)
            d.pop_in = None
            out[nbr] = d
        return out

    def _neighbour_block_descs(self, handles):
        """Block counts of each neighbour and where my faces land among its ghost blocks."""
        g = self.halo.field.grid
        out = {}
        for nbr in self.peer:
            n_blocks, n_alloc, n_ghost_down, n_down, n_up = handles[nbr][2]
            d = g.desc(None, None, None).clone()
            d.n_blocks, d.n_blocks_alloc, d.n_down, d.n_up = n_blocks, n_alloc, n_down, n_up
            # pushing up lands in the upper neighbour's ghost-DOWN blocks (they come first), pushing down in the lower
            # neighbour's ghost-UP blocks (after its ghost-down ones)
            first_ghost = n_blocks if nbr > g.part else n_blocks + n_ghost_down
            out[nbr] = (d, first_ghost)
        return out

    def push(self, streamIdx: int) -> None:
        """Stores my boundary planes into the neighbours' ghost planes and publishes the update counter there."""
        h = self.halo
        f = h.field
        g = f.grid
        bk = g.backend
        st = bk.streamHandle(streamIdx)
        mine = g.desc(f, None, None)
        dn, up = g.neighbours()
        self.count += 1
        k = self.count
        args = (f.elem_bytes, f.cardinality, h.lattice_q)
        if not self.block:
            # both faces and both signals in one launch
            pu = self.peer.get(up, (None, None)) if up is not None else (None, None)
            pd = self.peer.get(dn, (None, None)) if dn is not None else (None, None)
            capi.call("nlbm_dense_halo_push2", C.byref(mine), f.data.data_ptr(),
                      pu[0], g.sizes[up] if up is not None else 0, (pu[1] + 4 * FROM_BELOW) if up is not None else None,
                      pd[0], g.sizes[dn] if dn is not None else 0, (pd[1] + 4 * FROM_ABOVE) if dn is not None else None,
                      self.counter.data_ptr(), k, *args, st)
            return
        for nbr, direction, slot in ((up, +1, FROM_BELOW), (dn, -1, FROM_ABOVE)):
            if nbr is None:
                continue
            pf, pg = self.peer[nbr]
            dd, first_ghost = self._descs[nbr]
            capi.call("nlbm_block_halo_push", C.byref(mine), f.data.data_ptr(), C.byref(dd), pf, first_ghost, *args, direction, st)
            capi.call("nlbm_flag_signal", pg + 4 * slot, k, st)

    def wait(self, streamIdx: int) -> None:
        """Holds the stream until both neighbours' pushes of the matching update arrived in my ghost planes."""
        g = self.halo.field.grid
        dn, up = g.neighbours()
        self.waited += 1
        if dn is None and up is None:
            return
        base = self.flags.data_ptr()
        capi.call("nlbm_flag_wait2", (base + 4 * FROM_BELOW) if dn is not None else None, (base + 4 * FROM_ABOVE) if up is not None else None,
                  self.waited, TIMEOUT_MS, self.err.data_ptr(), g.backend.streamHandle(streamIdx))

    def run(self, streamIdx: int) -> None:
        """A complete halo update: push, then wait for the neighbours' pushes."""
        self.push(streamIdx)
        self.wait(streamIdx)

    def timeouts(self) -> int:
        return int(self.err.item())


class FusedIteration:
    """LBM iteration whose halo update is fused into the step kernel (nlbm_dense_step_push, include/neon_lbm.h).

    Per iteration and rank: [nlbm_flag_wait per neighbour] + ONE kernel.  The kernel takes the two z-boundary planes
    first, stores their face-crossing populations straight into the neighbours' ghost planes (CUDA-IPC mappings, NVLink
    stores) and publishes the iteration counter in the neighbours' flag words; the neighbours wait for that counter
    before their next launch.  Compared with Skeleton + Occ::standard (INTERNAL kernel next to halo + BOUNDARY kernel,
    libNeonSkeleton/src/skeleton/internal/multiGpuGraph.cpp:120-143) there is no view split, no copy kernel and no
    second stream.  Hazards: RAW — iteration t reads ghost planes written during the neighbours' iteration t-1 and waits
    for their counter t; WAR — a neighbour's ghost plane of field B is overwritten during my iteration t only after its
    boundary planes of iteration t-1 (the last readers) reported in, which is what its counter t says.
    """

    KIND = {(19, "float32", "float32"): 0, (19, "float64", "float64"): 1, (19, "float32", "float64"): 2,
            (27, "float32", "float32"): 3, (27, "float64", "float64"): 4}

    def __init__(self, pops, flag, omega: float, lattice_q: int, compute, arith: int, opts: int):
        import numpy as np
        self.pop, self.flag, self.omega, self.q = pops, flag, float(omega), lattice_q
        f = pops[0]
        g = f.grid
        bk = g.backend
        if getattr(g, "kind", "dense") != "dense" or bk.world < 2 or g.nz_local < 2:
            raise capi.NeonException("FusedIteration", capi.ERR_UNSUPPORTED,
                                     "the fused step + halo kernel needs a z-split dense grid with >= 2 planes per rank")
        cdt = np.dtype(compute).name if compute is not None else f.dtype.name
        self.kind = self.KIND[(lattice_q, f.dtype.name, cdt)]
        self.opts = int(arith) | int(opts)
        self.t = 0
        self.flags = torch.zeros(_FLAG_WORDS, dtype=torch.int32, device=bk.device)
        self.counters = torch.zeros(2, dtype=torch.int32, device=bk.device)
        self.err = torch.zeros(1, dtype=torch.int32, device=bk.device)
        torch.cuda.synchronize(bk.device)
        mine = (export_ptr(pops[0].data.data_ptr()), export_ptr(pops[1].data.data_ptr()), export_ptr(self.flags.data_ptr()))
        handles = [None] * bk.world
        dist.all_gather_object(handles, mine, group=bk.group)
        self.dn, self.up = g.neighbours()
        self.peer = {}
        for nbr in (self.dn, self.up):
            if nbr is not None:
                h0, h1, hf = handles[nbr]
                self.peer[nbr] = ((import_ptr(*h0), import_ptr(*h1)), import_ptr(*hf))
        # the ghost planes of the first input field must be current: one ordinary halo update (it also lines the ranks up)
        from .dgrid import StencilSemantic, TransferMode
        self._prologue = pops[0].newHaloUpdate(StencilSemantic.streaming, TransferMode.get, lattice_q, "ipc")
        bk.barrier()

    def run(self) -> None:
        g = self.pop[0].grid
        bk = g.backend
        st = bk.streamHandle(0)
        t = self.t
        fin, fout = self.pop[t & 1], self.pop[(t + 1) & 1]
        if t == 0:
            self._prologue.run(0)
        else:
            for nbr, slot in ((self.dn, FROM_BELOW), (self.up, FROM_ABOVE)):
                if nbr is not None:
                    capi.call("nlbm_flag_wait", self.flags.data_ptr() + 4 * slot, t, TIMEOUT_MS, self.err.data_ptr(), st)
        p = capi.PeerDesc()
        if self.dn is not None:
            fields, fl = self.peer[self.dn]
            p.down_field, p.down_nz_local, p.down_flag = fields[(t + 1) & 1], g.sizes[self.dn], fl + 4 * FROM_ABOVE
        if self.up is not None:
            fields, fl = self.peer[self.up]
            p.up_field, p.up_nz_local, p.up_flag = fields[(t + 1) & 1], g.sizes[self.up], fl + 4 * FROM_BELOW
        p.counters, p.value = self.counters.data_ptr(), t + 1
        d = g.desc(fin, fout, self.flag)
        capi.call("nlbm_dense_step_push", self.kind, C.byref(d), C.byref(p), self.omega, self.opts, st)
        self.t = t + 1

    def timeouts(self) -> int:
        return int(self.err.item()) + self._prologue.timeouts()
