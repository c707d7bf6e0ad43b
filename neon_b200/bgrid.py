"""bGrid / bField — block-sparse grid of 8 x 8 x 8-cell blocks, z-partitioned by block layers.

Mirrors libNeonDomain/include/Neon/domain/details/bGrid/ (Neon::bGrid = StaticBlock<8,8,8>, domain/bGrid.h:5):
  * blocks          bGrid_imp.h:7-185   a block exists when any of its cells is active (activeCellLambda)
  * connectivity    bGrid_imp.h:140-185, bPartition_imp.h:194-198   27 neighbour ids per block, index (dx+1)+3(dy+1)+9(dz+1)
  * active mask     StaticBlock.h:47-103   (here folded into the flag word: cells that are not active carry class UNDEFINED)
  * partitioning    tools/partitioning/SpanDecomposition.h:98-150   1-D over z, boundary blocks = first / last block layer
  * halo update     bField_imp.h:173-332   (upstream ignores the cardinality — NaN for Q = 19; here only the
                    facing z-slice of the crossing populations of every boundary block moves, nlbm_block_halo_push)

Layout: pop[q][blk][z][y][x] (include/neon_lbm.h, nlbm_block_desc).  One process per GPU: partition i lives on rank i.
Storage is flat torch tensors (device-memory plumbing); all arithmetic on them happens in libneon_lbm.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.distributed as dist

from . import _capi as capi
from .backend import Backend, Runtime
from .dgrid import _TORCH_DT, StencilSemantic, TransferMode, _aligned_zeros, partition_z

B = 8
BLOCK_CELLS = B * B * B
NO_BLOCK = 0xFFFFFFFF


def _ceil_div(a: int, b: int) -> int:
    return (a + b - 1) // b


class bGrid:
    kind = "block"

    def __init__(self, backend: Backend, dim: Sequence[int],
                 active: Union[None, np.ndarray, Callable[[np.ndarray, np.ndarray, np.ndarray], np.ndarray]] = None,
                 partition: Optional[Tuple[int, int]] = None):
        """``active``: None (every cell of the box), a bool array [nz, ny, nx], or a vectorised predicate
        active(x, y, z) -> bool array (the reference's activeCellLambda, bGrid_imp.h:7-30)."""
        self.backend = backend
        self.dim = tuple(int(v) for v in dim)
        nx, ny, nz = self.dim
        self.nb = (_ceil_div(nx, B), _ceil_div(ny, B), _ceil_div(nz, B))  # blocks per axis (x, y, z)
        nbx, nby, nbz = self.nb
        self.part, self.nparts = partition if partition is not None else (backend.rank, backend.world)

        # ---- which cells / blocks are active
        if active is None:
            act = None
            blk_active = np.ones((nbz, nby, nbx), bool)
        else:
            if callable(active):
                z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij", sparse=True)
                active = np.broadcast_to(active(x, y, z), (nz, ny, nx))
            act = np.zeros((nbz * B, nby * B, nbx * B), bool)
            act[:nz, :ny, :nx] = np.asarray(active, bool)
            blk_active = act.reshape(nbz, B, nby, B, nbx, B).any(axis=(1, 3, 5))
        self._cell_active = act

        # ---- z-partition by block layers (x, y whole), as dGrid does by planes
        sizes, origins = partition_z(nbz, self.nparts)
        if self.nparts > 1 and min(sizes) < 2:
            raise ValueError("every partition needs at least two block layers")
        l0, l1 = origins[self.part], origins[self.part] + sizes[self.part]
        self.layers = (l0, l1)

        def layer_blocks(lz: int) -> np.ndarray:
            by, bx = np.nonzero(blk_active[lz])
            return np.stack([np.full_like(bx, lz), by, bx], axis=1)  # rows (bz, by, bx), sorted by (by, bx)

        local = [layer_blocks(lz) for lz in range(l0, l1)]
        split = self.nparts > 1
        self.n_down = len(local[0]) if split and self.part > 0 else 0
        self.n_up = len(local[-1]) if split and self.part < self.nparts - 1 else 0
        ghost_dn = layer_blocks(l0 - 1) if split and self.part > 0 else np.zeros((0, 3), np.int64)
        ghost_up = layer_blocks(l1) if split and self.part < self.nparts - 1 else np.zeros((0, 3), np.int64)
        coords = np.concatenate(local + [ghost_dn, ghost_up], axis=0).astype(np.int64)
        self.n_blocks = int(sum(len(b) for b in local))
        self.n_ghost_down, self.n_ghost_up = len(ghost_dn), len(ghost_up)
        self.n_blocks_alloc = len(coords)
        self.block_coords = coords  # [n_alloc, 3] (bz, by, bx)
        if self.n_blocks_alloc >= NO_BLOCK:
            raise ValueError("too many blocks for 32-bit block ids")

        # ---- connectivity + origin: one 128-byte info line per block
        lut = np.full((nbz + 2, nby + 2, nbx + 2), NO_BLOCK, np.int64)
        lut[coords[:, 0] + 1, coords[:, 1] + 1, coords[:, 2] + 1] = np.arange(self.n_blocks_alloc)
        info = np.full((max(self.n_blocks_alloc, 1), 32), 0, np.uint32)
        info[:, :27] = NO_BLOCK
        lc = coords[:self.n_blocks]
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    k = (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)
                    info[:self.n_blocks, k] = lut[lc[:, 0] + 1 + dz, lc[:, 1] + 1 + dy, lc[:, 2] + 1 + dx].astype(np.uint32)
        info[:self.n_blocks_alloc, 27] = coords[:, 2] * B
        info[:self.n_blocks_alloc, 28] = coords[:, 1] * B
        info[:self.n_blocks_alloc, 29] = coords[:, 0] * B
        self.info_host = info
        self.info = _aligned_zeros(info.size, torch.int32, backend.device).view(-1, 32)
        self.info.copy_(torch.from_numpy(info.view(np.int32)))

        # ---- active-cell bit mask per block (bit z*64 + y*8 + x), only when the grid is not fully active
        self.active_mask = None
        if act is not None:
            cells = self._gather_blocks(act[None].astype(np.uint8))[0].astype(bool)  # [n_alloc, 512]
            words = np.packbits(cells.reshape(-1, 16, 32), axis=2, bitorder="little").view(np.uint32).reshape(-1, 16)
            self.active_mask = _aligned_zeros(max(words.size, 1), torch.int32, backend.device)
            self.active_mask[:words.size].copy_(torch.from_numpy(words.reshape(-1).view(np.int32)))
        self._field_uid = 0

    # --- dense <-> blocks ---------------------------------------------------------------------------------------------
    def _gather_blocks(self, glob: np.ndarray) -> np.ndarray:
        """[c, nz, ny, nx] (or padded to whole blocks) -> [c, n_alloc, 512] for this partition's local + ghost blocks."""
        nbx, nby, nbz = self.nb
        c = glob.shape[0]
        if glob.shape[1:] != (nbz * B, nby * B, nbx * B):
            pad = np.zeros((c, nbz * B, nby * B, nbx * B), glob.dtype)
            pad[:, :glob.shape[1], :glob.shape[2], :glob.shape[3]] = glob
            glob = pad
        v = glob.reshape(c, nbz, B, nby, B, nbx, B).transpose(0, 1, 3, 5, 2, 4, 6)  # [c, bz, by, bx, z, y, x]
        bc = self.block_coords
        return np.ascontiguousarray(v[:, bc[:, 0], bc[:, 1], bc[:, 2]]).reshape(c, self.n_blocks_alloc, BLOCK_CELLS)

    def _scatter_blocks(self, blocks: np.ndarray, fill=0) -> np.ndarray:
        """[c, n_blocks, 512] of the LOCAL blocks -> global dense [c, nz, ny, nx] (``fill`` elsewhere)."""
        nbx, nby, nbz = self.nb
        nx, ny, nz = self.dim
        c = blocks.shape[0]
        out = np.full((c, nbz, nby, nbx, B, B, B), fill, blocks.dtype)
        bc = self.block_coords[:self.n_blocks]
        out[:, bc[:, 0], bc[:, 1], bc[:, 2]] = blocks[:, :self.n_blocks].reshape(c, self.n_blocks, B, B, B)
        return np.ascontiguousarray(out.transpose(0, 1, 4, 2, 5, 3, 6).reshape(c, nbz * B, nby * B, nbx * B)[:, :nz, :ny, :nx])

    # --- descriptors / factories ----------------------------------------------------------------------------------------
    def desc(self, pop_in: Optional["bField"], pop_out: Optional["bField"], flag: Optional["bFlagField"]) -> capi.BlockDesc:
        d = capi.BlockDesc()
        d.pop_in = pop_in.data.data_ptr() if pop_in is not None else None
        d.pop_out = pop_out.data.data_ptr() if pop_out is not None else None
        d.flags = flag.words.data_ptr() if flag is not None else None
        d.info = self.info.data_ptr()
        d.n_blocks, d.n_blocks_alloc, d.n_down, d.n_up = self.n_blocks, self.n_blocks_alloc, self.n_down, self.n_up
        d.gnx, d.gny, d.gnz = self.dim
        return d

    def newField(self, name: str, cardinality: int, dtype=np.float32) -> "bField":
        self._field_uid += 1
        return bField(self, name, cardinality, np.dtype(dtype), self._field_uid)

    def newFlagField(self, name: str = "flag", like: Optional["bField"] = None) -> "bFlagField":
        self._field_uid += 1
        return bFlagField(self, name, self._field_uid)

    def getNumActiveCells(self) -> int:
        nx, ny, nz = self.dim
        return int(self._cell_active[:nz, :ny, :nx].sum()) if self._cell_active is not None else nx * ny * nz

    def neighbours(self) -> Tuple[Optional[int], Optional[int]]:
        r, n = self.part, self.nparts
        return (r - 1 if r > 0 else None), (r + 1 if r < n - 1 else None)


class _BlockFieldBase:
    def _merge(self, loc: np.ndarray) -> np.ndarray:
        """Global dense array from every rank's local blocks (test / validation helper)."""
        g = self.grid
        if g.backend.world == 1:
            return loc
        parts = [None] * g.backend.world
        dist.all_gather_object(parts, (g.block_coords[:g.n_blocks], loc), group=g.backend.group)
        out = loc.copy()
        nx, ny, nz = g.dim
        for bc, other in parts:
            zs = np.unique(bc[:, 0])
            for lz in zs:  # z-partition: whole block layers belong to one rank
                z0, z1 = lz * B, min((lz + 1) * B, nz)
                out[:, z0:z1] = other[:, z0:z1]
        return out


class bField(_BlockFieldBase):
    """Population (or any floating point) field over the blocks of a bGrid: ``cardinality`` SoA components."""

    def __init__(self, grid: bGrid, name: str, cardinality: int, dtype: np.dtype, uid: int):
        if dtype not in _TORCH_DT:
            raise TypeError(f"unsupported field type {dtype}")
        self.grid, self.name, self.cardinality, self.dtype, self.uid = grid, name, cardinality, dtype, uid
        self.elem_bytes = dtype.itemsize
        n = cardinality * max(grid.n_blocks_alloc, 1) * BLOCK_CELLS
        self.data = _aligned_zeros(n, _TORCH_DT[dtype], grid.backend.device)
        self.view3 = self.data.view(cardinality, max(grid.n_blocks_alloc, 1), BLOCK_CELLS)

    def updateDeviceData(self, host: np.ndarray, stream_idx: int = 0) -> None:
        """``host``: the GLOBAL dense field [cardinality, nz, ny, nx]; local and ghost blocks are filled from it."""
        g = self.grid
        nx, ny, nz = g.dim
        assert tuple(host.shape) == (self.cardinality, nz, ny, nx), host.shape
        blocks = g._gather_blocks(np.asarray(host, self.dtype))
        self.view3[:, :g.n_blocks_alloc].copy_(torch.from_numpy(blocks), non_blocking=False)

    # host mirror kept in the field's own layout [cardinality, blocks, 512] (what Neon's bField host mirror is): no
    # re-ordering on either side of the bus
    def updateDeviceBlocks(self, host: torch.Tensor, stream_idx: int = 0) -> None:
        """``host`` (pinned for an asynchronous copy): [cardinality, n_blocks_alloc, 512], local blocks then ghost blocks."""
        g = self.grid
        assert tuple(host.shape) == (self.cardinality, g.n_blocks_alloc, BLOCK_CELLS), host.shape
        with torch.cuda.stream(g.backend.stream(stream_idx)):
            self.view3[:, :g.n_blocks_alloc].copy_(host, non_blocking=True)

    def updateHostBlocksInto(self, host: torch.Tensor, stream_idx: int = 0) -> None:
        """Asynchronous device -> host copy of the LOCAL blocks into ``host[:, :n_blocks]``."""
        g = self.grid
        with torch.cuda.stream(g.backend.stream(stream_idx)):
            host[:, :g.n_blocks].copy_(self.view3[:, :g.n_blocks], non_blocking=True)

    def copyFrom(self, other: "bField", stream_idx: int = 0) -> None:
        """Device-to-device copy of another field of the same grid and shape (ghost blocks included)."""
        assert other.grid is self.grid and other.cardinality == self.cardinality and other.dtype == self.dtype
        with torch.cuda.stream(self.grid.backend.stream(stream_idx)):
            self.data.copy_(other.data, non_blocking=True)

    def updateHostData(self) -> np.ndarray:
        """Global dense array [cardinality, nz, ny, nx] holding this rank's local blocks (zeros elsewhere)."""
        g = self.grid
        return g._scatter_blocks(self.view3[:, :g.n_blocks].cpu().numpy())

    def gather(self) -> np.ndarray:
        return self._merge(self.updateHostData())

    def newHaloUpdate(self, semantic: StencilSemantic = StencilSemantic.standard, transfer: TransferMode = TransferMode.get,
                      lattice_q: int = 0, transport: str = "auto"):
        from .halo import HaloUpdateContainer
        return HaloUpdateContainer(self, semantic, transfer, lattice_q, transport)


class bFlagField(_BlockFieldBase):
    """Per-cell flag words over the blocks (class + wallNghBitflag); cells that are not active are UNDEFINED."""

    def __init__(self, grid: bGrid, name: str, uid: int):
        self.grid, self.name, self.uid = grid, name, uid
        self.cardinality, self.elem_bytes = 1, 4
        self.words = _aligned_zeros(max(grid.n_blocks_alloc, 1) * BLOCK_CELLS, torch.int32, grid.backend.device)
        self.cells = self.words.view(max(grid.n_blocks_alloc, 1), BLOCK_CELLS)

    def _d(self) -> capi.BlockDesc:
        return self.grid.desc(None, None, self)

    def setClasses(self, cls_global: np.ndarray, stream_idx: int = 0) -> None:
        """Upload cell classes [nz, ny, nx]; cells outside the box or not active become ``undefined``."""
        g = self.grid
        nx, ny, nz = g.dim
        assert cls_global.shape == (nz, ny, nx)
        nbx, nby, nbz = g.nb
        full = np.full((1, nbz * B, nby * B, nbx * B), capi.UNDEFINED, np.uint32)
        full[0, :nz, :ny, :nx] = cls_global.astype(np.uint32)
        if g._cell_active is not None:
            full[0][~g._cell_active] = capi.UNDEFINED
        blocks = g._gather_blocks(full)[0] << capi.FLAG_CLASS_SHIFT
        self.cells[:g.n_blocks_alloc].copy_(torch.from_numpy(blocks.astype(np.uint32).view(np.int32)))

    def classify(self, geom: int, sphere: Optional[Sequence[float]] = None, stream_idx: int = 0) -> None:
        g = self.grid
        sp = (C.c_double * 4)(*sphere) if sphere is not None else None
        am = g.active_mask.data_ptr() if g.active_mask is not None else None
        capi.call("nlbm_block_classify", C.byref(self._d()), geom, sp, am, g.backend.streamHandle(stream_idx))

    def computeWallNghMask(self, q: int, stream_idx: int = 0) -> None:
        """LbmContainers::computeWallNghMask (LbmTools.h:344-376) through the block connectivity, in place."""
        bk = self.grid.backend
        bad = torch.zeros(1, dtype=torch.int32, device=bk.device)
        capi.call("nlbm_block_wall_mask", C.byref(self._d()), q, bad.data_ptr(), bk.streamHandle(stream_idx))
        bk.sync(stream_idx)
        if int(bad.item()) != 0:
            raise capi.NeonException("computeWallNghMask", capi.ERR_GEOMETRY,
                                     f"{int(bad.item())} bulk-cell neighbours are missing or not active")

    def _local(self) -> np.ndarray:
        g = self.grid
        w = self.cells[:g.n_blocks].cpu().numpy().view(np.uint32)
        return g._scatter_blocks(w[None], fill=np.uint32(capi.UNDEFINED << capi.FLAG_CLASS_SHIFT))[0]

    def classes(self) -> np.ndarray:
        return ((self._local() >> capi.FLAG_CLASS_SHIFT) & 3).astype(np.int32)

    def masks(self) -> np.ndarray:
        return self._local() & capi.FLAG_MASK_BITS

    def gather(self):
        w = self._merge(self._local()[None])[0]
        return ((w >> capi.FLAG_CLASS_SHIFT) & 3).astype(np.int32), w & capi.FLAG_MASK_BITS
