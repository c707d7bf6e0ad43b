"""dGrid / dField — dense grid, z-slab partitioned, structure-of-arrays fields with a 512-byte aligned row pitch.

Mirrors libNeonDomain/include/Neon/domain/details/dGrid/:
  * partitioning   dGrid_imp.h:32-63   (floor(Z/n) planes each, the first Z mod n devices take one more; x, y whole)
  * halo radius    dGrid_imp.h:65-71   (stencil radius 1 -> one ghost plane per side when there is more than one device)
  * field layout   dField_imp.h:67-87  (reference: unpadded SoA; here pop[q][zm][y][x], pitch from nlbm_dense_layout)
  * halo update    dField_imp.h:341-421, 548-641 (newHaloUpdate)

One process per GPU: partition i lives on rank i of the Backend.  Storage is a flat torch tensor (device memory
plumbing); all arithmetic on it happens in libneon_lbm.so.
"""
from __future__ import annotations

import ctypes as C
from enum import Enum
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _capi as capi
from .backend import Backend, Runtime
from .lattice import crossing


class DataView(Enum):
    """Neon::DataView (libNeonCore/include/Neon/core/types/DataView.h:7-12)"""
    STANDARD = 0
    INTERNAL = 1
    BOUNDARY = 2


class StencilSemantic(Enum):
    """Neon::set::StencilSemantic: ``standard`` ("grid": every component crosses the face, --huGrid) or ``streaming``
    ("lattice": only populations whose c_z points across the face, --huLattice; throws in the reference with more
    than one device, dField_imp.h:610-612 — implemented here)."""
    standard = "standard"
    streaming = "streaming"


class TransferMode(Enum):
    """Neon::set::TransferMode (put: the source enqueues the copy, get: the destination does)."""
    put = "put"
    get = "get"


def partition_z(nz: int, n: int) -> Tuple[List[int], List[int]]:
    """dGrid_imp.h:43-62"""
    if n < 1 or nz < n:
        raise ValueError(f"cannot split {nz} planes over {n} device(s)")
    base, rem = divmod(nz, n)
    sizes = [base + (1 if i < rem else 0) for i in range(n)]
    origins = [sum(sizes[:i]) for i in range(n)]
    return sizes, origins


_TORCH_DT = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


def _aligned_zeros(n: int, dtype: torch.dtype, device: torch.device, align: int = 512) -> torch.Tensor:
    """Zeroed flat tensor whose base address is ``align``-byte aligned (CUDA allocations already are; host ones not)."""
    if device.type == "cuda":
        return torch.zeros(n, dtype=dtype, device=device)
    item = torch.empty((), dtype=dtype).element_size()
    raw = torch.zeros(n + align // item, dtype=dtype)
    off = (-raw.data_ptr() % align) // item
    return raw[off:off + n]


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class dGrid:
    kind = "dense"

    def __init__(self, backend: Backend, dim: Sequence[int], stencil_radius: int = 1,
                 partition: Optional[Tuple[int, int]] = None):
        """``partition=(index, count)`` places an arbitrary slab of a ``count``-way split on THIS process's device —
        the counterpart of the reference's oversubscribed device lists ({0,0,0}: several partitions on one GPU,
        libNeonDomain/tests/domain-halos/src/runHelper.h:64-67).  Default: partition ``rank`` of ``world``."""
        if stencil_radius != 1:
            raise ValueError("the LBM path uses radius-1 stencils (D3Q19 / D3Q27)")
        self.backend = backend
        self.dim = tuple(int(v) for v in dim)
        nx, ny, nz = self.dim
        self.part, self.nparts = partition if partition is not None else (backend.rank, backend.world)
        self.sizes, self.origins = partition_z(nz, self.nparts)
        self.nz_local = self.sizes[self.part]
        self.z_origin = self.origins[self.part]
        self.z_halo = 1 if self.nparts > 1 else 0  # dField_imp.h:48-51
        self.nzm = self.nz_local + 2 * self.z_halo
        self._field_uid = 0

    # --- layout -------------------------------------------------------------------------------------------------
    def layout(self, cardinality: int, elem_bytes: int):
        d = capi.DenseDesc()
        nx, ny, nz = self.dim
        d.nx, d.ny, d.nz_local, d.z_halo = nx, ny, self.nz_local, self.z_halo
        d.z_origin, d.gnx, d.gny, d.gnz = self.z_origin, nx, ny, nz
        pb, fb = C.c_size_t(), C.c_size_t()
        capi.call("nlbm_dense_layout", C.byref(d), cardinality, elem_bytes, C.byref(pb), C.byref(fb))
        return d, pb.value, fb.value

    def desc(self, pop_in: Optional["dField"], pop_out: Optional["dField"], flag: Optional["FlagField"]) -> capi.DenseDesc:
        ref = pop_in or pop_out
        d = (ref._desc if ref is not None else flag._desc).clone()
        d.pop_in = pop_in.data.data_ptr() if pop_in is not None else None
        d.pop_out = pop_out.data.data_ptr() if pop_out is not None else None
        d.flags = flag.words.data_ptr() if flag is not None else None
        d.wall_cache = pop_out.wallCachePtr() if isinstance(pop_out, dField) else None
        return d

    # --- factories (Grid::newField, dGrid.h) ---------------------------------------------------------------------
    def newField(self, name: str, cardinality: int, dtype=np.float32) -> "dField":
        self._field_uid += 1
        return dField(self, name, cardinality, np.dtype(dtype), self._field_uid)

    def newFlagField(self, name: str = "flag", like: Optional["dField"] = None) -> "FlagField":
        self._field_uid += 1
        return FlagField(self, name, 4 if like is None else like.elem_bytes, self._field_uid)

    def getNumActiveCells(self) -> int:
        return self.dim[0] * self.dim[1] * self.dim[2]

    def neighbours(self) -> Tuple[Optional[int], Optional[int]]:
        """(rank below, rank above): no periodic wrap (dField_imp.h:409-415)."""
        r, n = self.part, self.nparts
        return (r - 1 if r > 0 else None), (r + 1 if r < n - 1 else None)


class _FieldBase:
    def _global_planes(self):
        """(memory plane, global z) for every memory plane that lies inside the global box."""
        g = self.grid
        out = []
        for zm in range(g.nzm):
            gz = g.z_origin + zm - g.z_halo
            if 0 <= gz < g.dim[2]:
                out.append((zm, gz))
        return out


class dField(_FieldBase):
    """Population (or any floating point) field: ``cardinality`` SoA components over the local z-slab + ghost planes."""

    def __init__(self, grid: dGrid, name: str, cardinality: int, dtype: np.dtype, uid: int):
        if dtype not in _TORCH_DT:
            raise TypeError(f"unsupported field type {dtype}")
        self.grid, self.name, self.cardinality, self.dtype, self.uid = grid, name, cardinality, dtype, uid
        self.elem_bytes = dtype.itemsize
        self._desc, pop_bytes, _ = grid.layout(cardinality, self.elem_bytes)
        self.pitch_y, self.pitch_z, self.pitch_q = self._desc.pitch_y, self._desc.pitch_z, self._desc.pitch_q
        self.data = _aligned_zeros(pop_bytes // self.elem_bytes, _TORCH_DT[dtype], grid.backend.device)
        self.view4 = self.data.view(cardinality, grid.nzm, grid.dim[1], self.pitch_y)
        self._halo_buffers = {}
        # x-face cache (include/neon_lbm.h, nlbm_dense_wall_cache_build): valid from commitWalls() on; every writer of this
        # class refreshes it, code that pokes ``data`` directly calls commitWalls() or invalidateWalls() itself
        self._wall_cache = None
        self._walls_committed = False
        if grid.backend.runtime == Runtime.stream:
            nbytes = C.c_size_t()
            capi.call("nlbm_dense_wall_cache_layout", C.byref(self._desc), cardinality, self.elem_bytes, C.byref(nbytes))
            self._wall_cache = torch.zeros(nbytes.value // self.elem_bytes, dtype=_TORCH_DT[dtype], device=grid.backend.device)

    def commitWalls(self, stream_idx: int = 0) -> None:
        """(Re)builds the x-face cache from the field: call after writing the field other than through the step kernels."""
        if self._wall_cache is None:
            return
        d = self._desc.clone()
        d.pop_out, d.wall_cache = self.data.data_ptr(), self._wall_cache.data_ptr()
        capi.call("nlbm_dense_wall_cache_build", C.byref(d), self.cardinality, self.elem_bytes, self.grid.backend.streamHandle(stream_idx))
        self._walls_committed = True

    def invalidateWalls(self) -> None:
        self._walls_committed = False

    def wallCachePtr(self):
        return self._wall_cache.data_ptr() if self._walls_committed else None

    # --- host <-> device (FieldBase::updateDeviceData / updateHostData) -------------------------------------------
    def updateDeviceData(self, host, stream_idx: int = 0, host_z0: int = 0) -> None:
        """``host`` is the GLOBAL field [cardinality, nz, ny, nx] (numpy array or — for asynchronous copies — a pinned
        torch tensor); this rank takes its slab and, where they lie inside the box, its ghost planes.  Enqueued on the
        main stream like FieldBase::updateDeviceData(streamIdx).  A rank that holds only its own part of the host
        mirror (one process per GPU) passes the planes [host_z0, host_z0 + host.shape[1]) of the global field; they
        must cover its slab and in-box ghost planes."""
        g = self.grid
        nx, ny, nz = g.dim
        planes = self._global_planes()
        zm0, gz0 = planes[0]
        n = len(planes)
        assert tuple(host.shape[2:]) == (ny, nx) and host.shape[0] == self.cardinality, host.shape
        assert host_z0 <= gz0 and gz0 + n <= host_z0 + host.shape[1] <= nz, (host_z0, host.shape, gz0, n)
        if isinstance(host, np.ndarray):
            host = torch.from_numpy(np.ascontiguousarray(host, dtype=self.dtype))
        src = host[:, gz0 - host_z0:gz0 - host_z0 + n]
        if self.pitch_y == nx:  # rows are dense: one contiguous copy per component
            for q in range(self.cardinality):
                self.view4[q, zm0:zm0 + n].copy_(src[q], non_blocking=True)
        else:
            self.view4[:, zm0:zm0 + n, :, :nx].copy_(src, non_blocking=True)
        self.commitWalls(stream_idx)

    def copyFrom(self, other: "dField", stream_idx: int = 0) -> None:
        """Device-to-device copy of another field of the same grid and shape (ghost planes and x-face cache included):
        the second population field of the two-field scheme starts as a copy of the first, so a host mirror has to
        cross the bus only once."""
        assert other.grid is self.grid and other.cardinality == self.cardinality and other.dtype == self.dtype
        with torch.cuda.stream(self.grid.backend.stream(stream_idx)) if self.grid.backend.runtime == Runtime.stream else _nullctx():
            self.data.copy_(other.data, non_blocking=True)
            if self._wall_cache is not None and other._walls_committed:
                self._wall_cache.copy_(other._wall_cache, non_blocking=True)
                self._walls_committed = True
            else:
                self.commitWalls(stream_idx)

    def updateHostDataInto(self, host: torch.Tensor) -> None:
        """Asynchronous device -> host copy of the local slab into ``host`` (pinned, [cardinality, nz_local, ny, nx])."""
        g = self.grid
        nx = g.dim[0]
        if self.pitch_y == nx:
            for q in range(self.cardinality):
                host[q].copy_(self.view4[q, g.z_halo:g.z_halo + g.nz_local], non_blocking=True)
        else:
            host.copy_(self.view4[:, g.z_halo:g.z_halo + g.nz_local, :, :nx], non_blocking=True)

    def updateHostData(self) -> np.ndarray:
        """Local slab without ghosts/padding: [cardinality, nz_local, ny, nx]."""
        g = self.grid
        return self.view4[:, g.z_halo:g.z_halo + g.nz_local, :, :g.dim[0]].cpu().numpy()

    def gather(self) -> np.ndarray:
        """The global field on every rank (test / validation helper)."""
        loc = self.updateHostData()
        if self.grid.backend.world == 1:
            return loc
        parts = [None] * self.grid.backend.world
        dist.all_gather_object(parts, loc, group=self.grid.backend.group)
        return np.concatenate(parts, axis=1)

    def plane(self, q: int, zm: int) -> torch.Tensor:
        return self.view4[q, zm]

    # --- dField::newHaloUpdate (dField.h:84-87) -------------------------------------------------------------------
    def newHaloUpdate(self, semantic: StencilSemantic = StencilSemantic.standard, transfer: TransferMode = TransferMode.get,
                      lattice_q: int = 0, transport: str = "auto"):
        from .halo import HaloUpdateContainer
        return HaloUpdateContainer(self, semantic, transfer, lattice_q, transport)


class FlagField(_FieldBase):
    """Per-cell flag words (class + wallNghBitflag, include/neon_lbm.h) followed by the per-row chunk summary.
    Replaces the 8-byte CellType field of the reference (benchmarks/.../src/CellType.h:33-34)."""

    def __init__(self, grid: dGrid, name: str, pop_elem_bytes: int, uid: int):
        self.grid, self.name, self.uid = grid, name, uid
        self.cardinality, self.elem_bytes = 1, 4
        self._desc, _, flag_bytes = grid.layout(1, pop_elem_bytes)  # flags share the populations' pitch
        self.pitch_y, self.pitch_z = self._desc.pitch_y, self._desc.pitch_z
        self.words = _aligned_zeros(flag_bytes // 4, torch.int32, grid.backend.device)
        self.cells = self.words[: grid.nzm * self.pitch_z].view(grid.nzm, grid.dim[1], self.pitch_y)

    def _d(self) -> capi.DenseDesc:
        return self.grid.desc(None, None, self)

    def setClasses(self, cls_global, stream_idx: int = 0, host_z0: int = 0) -> None:
        """Upload cell classes [nz, ny, nx] (0 bounceBack, 1 movingWall, 2 bulk; any integer numpy array, or a pinned
        uint8 torch tensor for an asynchronous copy); padding and planes outside the box become ``undefined``; wall
        bits are cleared.  ``host_z0``: global z of the first plane of ``cls_global`` when the caller holds only the
        planes this rank needs (slab + in-box ghost planes).  One byte per cell crosses the bus; the flag words are
        made on the device (nlbm_dense_flags_from_classes)."""
        g = self.grid
        nx, ny, nz = g.dim
        assert tuple(cls_global.shape[1:]) == (ny, nx) and host_z0 + cls_global.shape[0] <= nz, cls_global.shape
        planes = self._global_planes()  # consecutive memory planes <-> consecutive global planes
        (zm0, gz0), n = planes[0], len(planes)
        part = cls_global[gz0 - host_z0:gz0 - host_z0 + n]
        if g.backend.runtime != Runtime.stream:  # host-logic runtime: flag words written directly
            host = np.full((g.nzm, ny, self.pitch_y), capi.UNDEFINED << capi.FLAG_CLASS_SHIFT, np.uint32)
            np.left_shift(np.asarray(part), capi.FLAG_CLASS_SHIFT, out=host[zm0:zm0 + n, :, :nx], casting="unsafe")
            self.cells.copy_(torch.from_numpy(host.view(np.int32)))
            return
        if isinstance(part, np.ndarray):
            part = torch.from_numpy(np.ascontiguousarray(part, dtype=np.uint8))
        assert part.dtype == torch.uint8
        bk = g.backend
        with torch.cuda.stream(bk.stream(stream_idx)):
            dev = part.to(bk.device, non_blocking=True)
            capi.call("nlbm_dense_flags_from_classes", C.byref(self._d()), dev.data_ptr(), zm0, n, bk.streamHandle(stream_idx))
            dev.record_stream(bk.stream(stream_idx))

    def classify(self, geom: int, sphere: Optional[Sequence[float]] = None, stream_idx: int = 0) -> None:
        """Device-side geometry (RunCavityTwoPop.cu:208-224; SURVEY.md §8d for the sphere cases)."""
        sp = (C.c_double * 4)(*sphere) if sphere is not None else None
        capi.call("nlbm_dense_classify", C.byref(self._d()), geom, sp, self.grid.backend.streamHandle(stream_idx))

    def computeWallNghMask(self, q: int, stream_idx: int = 0) -> None:
        """LbmContainers::computeWallNghMask (LbmTools.h:344-376), run in place as RunCavityTwoPop.cu:239 does.
        Raises if a bulk cell has a neighbour outside the box (the reference would read invalid data)."""
        bk = self.grid.backend
        bad = torch.zeros(1, dtype=torch.int32, device=bk.device)
        capi.call("nlbm_dense_wall_mask", C.byref(self._d()), q, bad.data_ptr(), bk.streamHandle(stream_idx))
        bk.sync(stream_idx)
        if int(bad.item()) != 0:
            raise capi.NeonException("computeWallNghMask", capi.ERR_GEOMETRY,
                                     f"{int(bad.item())} bulk-cell neighbours fall outside the domain")

    def _local(self) -> np.ndarray:
        g = self.grid
        w = self.cells[g.z_halo:g.z_halo + g.nz_local, :, :g.dim[0]].cpu().numpy().view(np.uint32)
        return w

    def classes(self) -> np.ndarray:
        return ((self._local() >> capi.FLAG_CLASS_SHIFT) & 3).astype(np.int32)

    def masks(self) -> np.ndarray:
        return self._local() & capi.FLAG_MASK_BITS

    def gather(self):
        cls, msk = self.classes(), self.masks()
        if self.grid.backend.world == 1:
            return cls, msk
        parts = [None] * self.grid.backend.world
        dist.all_gather_object(parts, (cls, msk), group=self.grid.backend.group)
        return np.concatenate([p[0] for p in parts], 0), np.concatenate([p[1] for p in parts], 0)
