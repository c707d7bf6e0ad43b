"""Container — the unit of work a Skeleton schedules.

Mirrors Neon::set::Container (libNeonSet/include/Neon/set/Containter.h:16-167): ``run(streamIdx, dataView)`` enqueues
the work on stream ``streamIdx`` of the Backend and returns (asynchronous, Containter.h:25-27).  The reference parses
the user's loading lambda into data tokens (Loader::load, container/Loader.h:67-82; READ/WRITE x MAP/STENCIL) from
which the Skeleton derives dependencies and halo updates; here every container states its tokens explicitly.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Callable, List, Optional

from .dgrid import DataView, StencilSemantic


class Access(Enum):
    READ = "r"
    WRITE = "w"


class Pattern(Enum):
    """Neon::Pattern (older spelling Neon::Compute, benchmarks/lbm-flow-over-sphere/src/LbmContainers.h:455-459)"""
    MAP = "map"
    STENCIL = "stencil"


@dataclass
class Token:
    field: object
    access: Access
    pattern: Pattern
    semantic: StencilSemantic = StencilSemantic.standard
    lattice_q: int = 0


class Container:
    def __init__(self, name: str, tokens: List[Token], launch: Callable[[int, DataView], None], kind: str = "compute"):
        self.name, self.tokens, self._launch, self.kind = name, tokens, launch, kind

    def run(self, streamIdx: int = 0, dataView: DataView = DataView.STANDARD) -> None:
        self._launch(streamIdx, dataView)

    def getName(self) -> str:
        return self.name

    def stencilReads(self) -> List[Token]:
        return [t for t in self.tokens if t.access == Access.READ and t.pattern == Pattern.STENCIL]
