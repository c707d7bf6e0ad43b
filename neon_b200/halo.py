"""Halo update of a z-slab partitioned dense field.

Reference: dField::newHaloUpdate + initHaloUpdateTable (libNeonDomain/.../dGrid/dField_imp.h:341-421, 548-641) build, per
device and direction, one MemoryTransfer per population (19 cudaMemcpyPeerAsync per direction,
DataTransferContainer.h:38-55) fenced by host-blocking stream syncs (SynchronizationContainer.h:37-42).

Here a halo update is two messages per neighbour, ordered on the issuing stream only (no host sync):
  upward   : my top boundary plane    -> lower ghost plane of rank+1   (populations with c_z = +1, or all components)
  downward : my bottom boundary plane -> upper ghost plane of rank-1   (populations with c_z = -1, or all components)
Transports:
  "packed" (CUDA): nlbm_dense_halo_pack -> NCCL send/recv (grouped) -> nlbm_dense_halo_unpack
  "views"  (any) : send/recv straight from/into the contiguous population planes (also the gloo/CPU host-logic path)
  "ipc"    (CUDA): nlbm_dense_halo_push stores straight into the neighbour's ghost plane through a CUDA-IPC mapping,
                   ordered by device-side flags (see ipc.py)
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.distributed as dist

from . import _capi as capi
from .backend import Runtime
from .containers import Access, Container, Pattern, Token
from .dgrid import DataView, StencilSemantic, TransferMode
from .lattice import crossing


class HaloUpdateContainer(Container):
    def __init__(self, field, semantic: StencilSemantic, transfer: TransferMode, lattice_q: int, transport: str = "auto"):
        self.field, self.semantic, self.transfer = field, semantic, transfer
        if semantic == StencilSemantic.streaming and lattice_q not in (19, 27):
            raise ValueError("the streaming (lattice) semantic needs the lattice: lattice_q = 19 or 27")
        self.lattice_q = lattice_q if semantic == StencilSemantic.streaming else 0
        bk = field.grid.backend
        if transport == "auto":
            # CUDA: peer stores through CUDA-IPC mappings (one node, NVLink/NVSwitch: every bench/test topology here);
            # set NEON_B200_HALO=packed|views to route the faces through NCCL instead
            import os
            transport = os.environ.get("NEON_B200_HALO", "ipc") if bk.runtime == Runtime.stream else "views"
        if transport not in ("packed", "views", "ipc"):
            raise ValueError(transport)
        if getattr(field.grid, "kind", "dense") == "block" and transport != "ipc":
            raise ValueError("block-sparse fields exchange their faces through the peer-store transport (\"ipc\") only")
        if transport != "views" and bk.runtime != Runtime.stream:
            raise ValueError(f"transport {transport!r} needs CUDA")
        self.transport = transport
        self._bufs = {}
        self._ipc = None
        super().__init__(f"haloUpdate({field.name},{semantic.value},{transfer.value},{transport})",
                         [Token(field, Access.WRITE, Pattern.MAP)], self._run, kind="halo")

    def components(self, direction: int) -> List[int]:
        if self.lattice_q:
            return crossing(self.lattice_q, direction)
        return list(range(self.field.cardinality))

    def bytesPerDirection(self, direction: int) -> int:
        g = self.field.grid
        if getattr(g, "kind", "dense") == "block":  # one z-slice of 64 cells per boundary block
            return len(self.components(direction)) * (g.n_up if direction > 0 else g.n_down) * 64 * self.field.elem_bytes
        return len(self.components(direction)) * self.field.pitch_z * self.field.elem_bytes

    # ------------------------------------------------------------------------------------------------------------
    def _run(self, streamIdx: int, dataView: DataView) -> None:
        g = self.field.grid
        bk = g.backend
        if bk.world == 1:
            return
        if self.transport == "views":
            self._run_views(streamIdx)
        elif self.transport == "packed":
            self._run_packed(streamIdx)
        else:
            self._run_ipc(streamIdx)

    def _p2p(self, ops, streamIdx):
        bk = self.field.grid.backend
        if not ops:
            return
        if bk.runtime == Runtime.stream:
            with torch.cuda.stream(bk.stream(streamIdx)):
                for r in dist.batch_isend_irecv(ops):
                    r.wait()  # stream-level wait: the host does not block
        else:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def _run_views(self, streamIdx: int) -> None:
        f, g = self.field, self.field.grid
        dn, up = g.neighbours()
        top, bot = g.z_halo + g.nz_local - 1, g.z_halo
        ghost_lo, ghost_hi = 0, g.z_halo + g.nz_local
        grp = g.backend.group
        ops = []
        if up is not None:
            ops += [dist.P2POp(dist.isend, f.plane(q, top), up, grp) for q in self.components(+1)]
            ops += [dist.P2POp(dist.irecv, f.plane(q, ghost_hi), up, grp) for q in self.components(-1)]
        if dn is not None:
            ops += [dist.P2POp(dist.isend, f.plane(q, bot), dn, grp) for q in self.components(-1)]
            ops += [dist.P2POp(dist.irecv, f.plane(q, ghost_lo), dn, grp) for q in self.components(+1)]
        self._p2p(ops, streamIdx)

    def _buf(self, key: str, direction: int) -> torch.Tensor:
        if key not in self._bufs:
            n = len(self.components(direction)) * self.field.pitch_z
            self._bufs[key] = torch.empty(n, dtype=self.field.data.dtype, device=self.field.data.device)
        return self._bufs[key]

    def _run_packed(self, streamIdx: int) -> None:
        f, g = self.field, self.field.grid
        bk = g.backend
        st = bk.streamHandle(streamIdx)
        d = g.desc(f, None, None)
        dn, up = g.neighbours()
        grp = bk.group
        args = (f.elem_bytes, f.cardinality, self.lattice_q)
        ops = []
        if up is not None:
            capi.call("nlbm_dense_halo_pack", C.byref(d), f.data.data_ptr(), *args, +1, self._buf("s_up", +1).data_ptr(), None, st)
            ops += [dist.P2POp(dist.isend, self._buf("s_up", +1), up, grp), dist.P2POp(dist.irecv, self._buf("r_up", -1), up, grp)]
        if dn is not None:
            capi.call("nlbm_dense_halo_pack", C.byref(d), f.data.data_ptr(), *args, -1, self._buf("s_dn", -1).data_ptr(), None, st)
            ops += [dist.P2POp(dist.isend, self._buf("s_dn", -1), dn, grp), dist.P2POp(dist.irecv, self._buf("r_dn", +1), dn, grp)]
        self._p2p(ops, streamIdx)
        if up is not None:  # what came down from above lands in my upper ghost plane
            capi.call("nlbm_dense_halo_unpack", C.byref(d), f.data.data_ptr(), *args, -1, self._buf("r_up", -1).data_ptr(), st)
        if dn is not None:
            capi.call("nlbm_dense_halo_unpack", C.byref(d), f.data.data_ptr(), *args, +1, self._buf("r_dn", +1).data_ptr(), st)

    def timeouts(self) -> int:
        """Waits of the peer-store transport that gave up (a neighbour never signalled); 0 for the other transports."""
        return 0 if self._ipc is None else self._ipc.timeouts()

    def _ipc_halo(self):
        from .ipc import IpcHalo
        if self._ipc is None:
            self._ipc = IpcHalo(self)
        return self._ipc

    def _run_ipc(self, streamIdx: int) -> None:
        self._ipc_halo().run(streamIdx)

    # the two halves of an update, for schedules that push a field right after it was written and wait right before it is
    # read (Skeleton with Options.pipelinedHalo); peer-store transport only
    def supportsSplit(self) -> bool:
        return self.transport == "ipc" and self.field.grid.backend.world > 1

    def push(self, streamIdx: int) -> None:
        self._ipc_halo().push(streamIdx)

    def wait(self, streamIdx: int) -> None:
        self._ipc_halo().wait(streamIdx)
