"""The LBM hot path behind the reference's names.

  LbmContainers.iteration          benchmarks/lbm-lid-driven-cavity-flow/src/LbmTools.h:285-325
  LbmContainers.computeWallNghMask LbmTools.h:344-376
  LbmContainers.computeRhoAndU     LbmTools.h:384-437
  LbmIteration                     src/LbmIteration.h:19-101 (two pre-built Skeletons, parity flip, no field swap)
  getLbmParameters                 src/Config.cpp:105-111

Every container is one call into libneon_lbm.so (hand-written sm_100a kernels); nothing here computes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _capi as capi
from .backend import Runtime
from .containers import Access, Container, Pattern, Token
from .dgrid import DataView, FlagField, StencilSemantic, TransferMode, dField
from .skeleton import Occ, Options, Skeleton


def omega_from_re(n: int, re: float = 100.0, ulb: float = 0.04) -> float:
    """Config::helpSetLbmParameters, Config.cpp:105-111 (clength = N - 2)"""
    nu = ulb * float(n - 2) / re
    return 1.0 / (3.0 * nu + 0.5)


def _step_symbol(q: int, store: np.dtype, compute: Optional[np.dtype], kind: str = "dense") -> str:
    store = np.dtype(store)
    compute = store if compute is None else np.dtype(compute)
    if kind == "block":
        if store == compute and store in (np.float32, np.float64):
            return f"nlbm_d3q{q}_{'f32' if store == np.float32 else 'f64'}_block_step"
        raise ValueError(f"unsupported bGrid D3Q{q} store/compute pair {store}/{compute}")
    if store == np.float32 and compute == np.float32:
        return f"nlbm_d3q{q}_f32_dense_step"
    if store == np.float64 and compute == np.float64:
        return f"nlbm_d3q{q}_f64_dense_step"
    if q == 19 and store == np.float32 and compute == np.float64:
        return "nlbm_d3q19_f32c64_dense_step"
    raise ValueError(f"unsupported D3Q{q} store/compute pair {store}/{compute}")


class LbmContainers:
    @staticmethod
    def iteration(stencilSemantic: StencilSemantic, fIn: dField, fOut: dField, cellTypeField: FlagField, omega: float,
                  lattice_q: int = 19, compute=None, arith: int = capi.ARITH_FAST, opts: int = 0,
                  halo_transport: str = "auto") -> Container:
        """One fused pull-stream + BGK collide over the cells of ``dataView``.  fIn is loaded as const STENCIL (needs
        its ghost planes current), fOut as MAP write, the flags as MAP read (LbmTools.h:296-299)."""
        if fIn is fOut:
            raise capi.NeonException("LbmContainers.iteration", capi.ERR_INVALID, "fIn and fOut must be two fields")
        grid = fIn.grid
        sym = _step_symbol(lattice_q, fIn.dtype, compute, getattr(grid, "kind", "dense"))
        if fIn.cardinality != lattice_q or fOut.cardinality != lattice_q or fOut.dtype != fIn.dtype:
            raise capi.NeonException("LbmContainers.iteration", capi.ERR_INVALID, "population fields do not match the lattice")
        desc = grid.desc(fIn, fOut, cellTypeField)
        fn = getattr(capi.lib(), sym)
        bk = grid.backend
        o = int(arith) | int(opts)

        dense = getattr(grid, "kind", "dense") == "dense"

        def launch(streamIdx: int, dataView: DataView) -> None:
            if dense:  # the output field's x-face cache, if it is current (dField.commitWalls)
                desc.wall_cache = fOut.wallCachePtr()
            capi.check(fn(C.byref(desc), omega, dataView.value, o, bk.streamHandle(streamIdx)), sym)

        c = Container(f"LBM_iteration_D3Q{lattice_q}",
                      [Token(fIn, Access.READ, Pattern.STENCIL, stencilSemantic, lattice_q),
                       Token(fOut, Access.WRITE, Pattern.MAP), Token(cellTypeField, Access.READ, Pattern.MAP)], launch)
        c.halo_transport = halo_transport
        # the next iteration stencil-reads fOut: a pipelined Skeleton pushes its faces right after the BOUNDARY kernel
        c.push_after = [Token(fOut, Access.WRITE, Pattern.STENCIL, stencilSemantic, lattice_q)]
        return c

    @staticmethod
    def computeWallNghMask(infoInField: FlagField, infoOutpeField: FlagField, lattice_q: int = 19) -> Container:
        if infoInField is not infoOutpeField:
            raise capi.NeonException("computeWallNghMask", capi.ERR_UNSUPPORTED,
                                     "the mask is built in place, as RunCavityTwoPop.cu:239 runs it")

        def launch(streamIdx: int, dataView: DataView) -> None:
            infoInField.computeWallNghMask(lattice_q, streamIdx)

        return Container("computeWallNghMask", [Token(infoInField, Access.READ, Pattern.STENCIL),
                                                Token(infoOutpeField, Access.WRITE, Pattern.MAP)], launch)

    @staticmethod
    def computeRhoAndU(fIn: dField, cellTypeField: FlagField, rho: dField, u: dField) -> Container:
        grid = fIn.grid
        sym = "nlbm_d3q19_f32_dense_rho_u" if fIn.dtype == np.float32 else "nlbm_d3q19_f64_dense_rho_u"
        if fIn.cardinality != 19 or rho.cardinality != 1 or u.cardinality != 3 or rho.dtype != fIn.dtype or u.dtype != fIn.dtype:
            raise capi.NeonException("computeRhoAndU", capi.ERR_INVALID, "rho must have 1 and u 3 components of fIn's type")
        desc = grid.desc(fIn, None, cellTypeField)
        fn = getattr(capi.lib(), sym)
        bk = grid.backend

        def launch(streamIdx: int, dataView: DataView) -> None:
            capi.check(fn(C.byref(desc), rho.data.data_ptr(), u.data.data_ptr(), bk.streamHandle(streamIdx)), sym)

        return Container("LBM_macroscopic", [Token(fIn, Access.READ, Pattern.STENCIL), Token(rho, Access.WRITE, Pattern.MAP),
                                             Token(u, Access.WRITE, Pattern.MAP)], launch)


class LbmIteration:
    """LbmIterationD3Q19 (LbmIteration.h:19-101), for D3Q19 and D3Q27: skeleton 0 streams pop0 -> pop1, skeleton 1
    pop1 -> pop0; run() executes skeleton[parity] and flips the parity."""

    def __init__(self, stencilSemantic: StencilSemantic, occ: Occ, transfer: TransferMode, fInField: dField,
                 fOutField: dField, flagField: FlagField, omega: float, lattice_q: int = 19, compute=None,
                 arith: int = capi.ARITH_FAST, opts: int = 0, halo_transport: str = "auto", graph: bool = False,
                 pipelined: bool = True):
        self.pop = [fInField, fOutField]
        self.flag, self.omega, self.parity = flagField, omega, 0
        self._step_args = (lattice_q, compute, int(arith) | int(opts))
        bk = fInField.grid.backend
        self.lbmTwoPop = []
        self.fused = None
        if halo_transport == "fused" and bk.world > 1:
            # B200-first alternative to the Skeleton's OCC graph: the halo update is part of the step kernel
            from .ipc import FusedIteration
            self.fused = FusedIteration(self.pop, flagField, omega, lattice_q, compute, arith, opts)
            return
        if halo_transport == "fused":
            halo_transport = "auto"
        for a, b in ((0, 1), (1, 0)):
            c = LbmContainers.iteration(stencilSemantic, self.pop[a], self.pop[b], flagField, omega, lattice_q, compute,
                                        arith, opts, halo_transport)
            sk = Skeleton(bk)
            # pipelined: iteration t pushes the faces of ITS output right after its BOUNDARY kernel, iteration t+1 only
            # waits for them (peer-store transport; other transports keep the update in front of the consumer)
            sk.sequence([c], f"LBM_{a}{b}", Options(occ, transfer, pipelinedHalo=pipelined), graph=graph)
            self.lbmTwoPop.append(sk)

    def run(self) -> None:
        if self.fused is not None:
            self.fused.run()
        else:
            self.lbmTwoPop[self.parity].run()
        self.parity ^= 1

    _KINDS = {(19, "float32", "float32"): 0, (19, "float64", "float64"): 1, (19, "float32", "float64"): 2,
              (27, "float32", "float32"): 3, (27, "float64", "float64"): 4}

    def runMany(self, iterations: int) -> None:
        """``iterations`` iterations with ONE library call (nlbm_dense_step_n) when the field lives on one device as a dense
        partition — the regime of small boxes, where an iteration lasts ~10 us as a kernel of its own: a chain of dependent
        launches whose first tiles wait plane-wise for the previous iteration (launch gap, ramp-up and tail of consecutive
        iterations overlap).
        Anything else (several partitions, bGrid) runs ``iterations`` x run()."""
        f = self.pop[self.parity]
        g = f.grid
        bk = g.backend
        q, compute, o = self._step_args
        if (iterations < 1 or bk.world > 1 or bk.runtime != Runtime.stream or self.fused is not None or getattr(g, "kind", "dense") != "dense"
                or g.z_halo != 0):
            for _ in range(iterations):
                self.run()
            return
        fin, fout = self.pop[self.parity], self.pop[self.parity ^ 1]
        kind = self._KINDS[(q, str(fin.dtype), str(fin.dtype if compute is None else np.dtype(compute)))]
        desc = g.desc(fin, fout, self.flag)
        capi.call("nlbm_dense_step_n", kind, C.byref(desc), fin.wallCachePtr(), self.omega, iterations, o, bk.streamHandle(0))
        self.parity ^= iterations & 1

    def chainPays(self) -> bool:
        """Whether runMany's launch chain beats a replay of single-iteration launches for this field: one dense partition of
        more than ~400 000 cells (below, an iteration is one chip-load of blocks and the chain is bound by the latency of one
        tile like the kernel itself: 64^3 22.2 against 24.1 GLUPS) up to 2^24 cells (256^3: +1.7 %; no difference above).
        Measured on B200, profiles/r02o_small_sweep.log: 80^3 +3.6 %, 96^3 +6 %, 128^3 +7.8 %, 160^3 +4 %, 192^3 +3.4 %."""
        g = self.pop[0].grid
        bk = g.backend
        if bk.world > 1 or bk.runtime != Runtime.stream or self.fused is not None or getattr(g, "kind", "dense") != "dense" or g.z_halo != 0:
            return False
        cells = g.dim[0] * g.dim[1] * g.dim[2]
        return 400_000 < cells <= (1 << 24)

    def runGraph(self, iterations: int, many: Optional[bool] = None) -> int:
        """ONE device: ``iterations`` (rounded up to an even count, so that the field parity is back where it started)
        captured once into a CUDA graph and replayed with one host call (``many``: one runMany(n) call — the launch chain —
        instead of n run() calls; default: whichever is faster for the box, chainPays()).  For boxes of a few hundred thousand cells one
        iteration takes ~10 us on a B200 — the host cannot issue launches that fast, and a launch costs as much as the
        kernel.  (The reference re-parses the loading lambda and calls cudaFuncGetAttributes for every launch,
        libNeonSys/include/Neon/sys/devices/gpu/GpuDevice.h:151-191.)  Returns the number of iterations actually run."""
        import torch
        bk = self.pop[0].grid.backend
        n = iterations + (iterations & 1)
        if bk.world > 1 or bk.runtime != Runtime.stream or self.fused is not None:
            for _ in range(n):
                self.run()
            return n
        if many is None:
            many = self.chainPays()
        cache = self.__dict__.setdefault("_graphs", {})
        key = (n, self.parity, many)
        main = bk.stream(0)
        if key not in cache:
            # make sure every lazily built object (x-face caches, function attributes, the chain's counters) exists before the capture
            if many:
                self.runMany(2)
            else:
                self.run()
                self.run()
            main.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(main):
                g.capture_begin()
                try:
                    if many:  # the launch chain of runMany (memset of its plane counters + n dependent launches) as one graph
                        self.runMany(n)
                    else:
                        for _ in range(n):
                            self.run()
                finally:
                    g.capture_end()
            cache[key] = g
            # the two un-captured runs above are part of the simulated time line: the caller asked for n, got n + 2 once
            with torch.cuda.stream(main):
                g.replay()
            return n + 2
        with torch.cuda.stream(main):
            cache[key].replay()
        return n

    def timeouts(self) -> int:
        """Device-side waits on a neighbour that gave up (peer-store transports); 0 when all faces arrived."""
        if self.fused is not None:
            return self.fused.timeouts()
        return sum(h.timeouts() for sk in self.lbmTwoPop for h in sk.halos())

    def getInput(self) -> dField:
        return self.pop[self.parity]

    def getOutput(self) -> dField:
        return self.pop[self.parity ^ 1]
