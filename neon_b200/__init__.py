"""neon_b200 — B200-native (sm_100a) implementation of Autodesk/Neon's Lattice-Boltzmann hot path.

Only what the path needs: the C-ABI kernel library (csrc/ -> lib/libneon_lbm.so, include/neon_lbm.h) and the host-side
mirror of the reference interface for it (Backend, dGrid/dField, Container, Skeleton with OCC, LbmIteration).
"""
from ._capi import (ARITH_FAST, ARITH_REFERENCE, BOUNCE_BACK, BULK, KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TMA, MOVING_WALL, UNDEFINED,
                    NeonException, OPT_FLAG_WORDS, OPT_REF_LITERAL, OPT_FLAGS_SUMMARY_FIRST, OPT_NO_XFACE_FIXUP_PREFETCH, OPT_NO_XFACE_PREFETCH,
                    opt_kernel, opt_rows_log2, opt_vec)
from .backend import Backend, Runtime
from .bgrid import bField, bFlagField, bGrid
from .containers import Access, Container, Pattern, Token
from .dgrid import DataView, FlagField, StencilSemantic, TransferMode, dField, dGrid, partition_z
from .lbm import LbmContainers, LbmIteration, omega_from_re
from .skeleton import Occ, Options, Skeleton

__all__ = [n for n in dir() if not n.startswith("_")]
