"""Backend — the device set, its stream set and event set, one PROCESS per GPU.

Mirrors Neon::Backend (libNeonSet/include/Neon/set/Backend.h:26-302).  The reference drives all GPUs of a node from
one process with one OpenMP host thread per device (DevSet.h:372-391); here every rank of a ``torch.distributed``
job owns exactly one device, so "device i of the backend" is rank i and ``devSet().setCardinality()`` is the world
size.  Streams and events are CUDA streams/events of that one device (torch wrappers: plumbing only).
"""
from __future__ import annotations

import os
from enum import Enum
from typing import List, Optional

import torch
import torch.distributed as dist


class Runtime(Enum):
    """Neon::Runtime (libNeonCore/include/Neon/core/types/Execution.h).  ``openmp`` is host logic only:
    fields live in host memory, halo exchange and scheduling work, compute containers refuse to run."""
    stream = "stream"
    openmp = "openmp"


class Backend:
    mainStreamIdx = 0  # Neon::Backend::mainStreamIdx

    def __init__(self, devices: Optional[List[int]] = None, runtime: Runtime = Runtime.stream, group=None):
        self.runtime = runtime
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world = dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        if devices is not None and len(devices) != self.world:
            raise ValueError(f"{len(devices)} device ids for a job of {self.world} rank(s): one process per GPU")
        if runtime == Runtime.stream:
            if not torch.cuda.is_available():
                raise RuntimeError("Runtime.stream needs a CUDA device (neon_b200 has no CPU compute path)")
            local = devices[self.rank] if devices is not None else int(os.environ.get("LOCAL_RANK", "0"))
            self.device = torch.device("cuda", local)
            torch.cuda.set_device(self.device)
        else:
            self.device = torch.device("cpu")
        self._streams: List[Optional[torch.cuda.Stream]] = []
        self.setAvailableStreamSet(1)

    # --- Backend.h:230-261
    def getDeviceCount(self) -> int:
        return self.world

    def setAvailableStreamSet(self, n: int) -> None:
        """Grows the stream set (the reference REPLACES it on the openmp runtime, Backend.cpp:363-366 — SURVEY fact 5)."""
        while len(self._streams) < n:
            if self.runtime == Runtime.stream:
                # stream 0 is the main stream (made torch's current stream so that tensor copies are ordered with
                # the kernels; an explicit stream, so a whole Skeleton run can be captured into a CUDA graph); the
                # others are side streams for halo + boundary work, created with HIGH priority: a small BOUNDARY kernel or
                # face copy must not queue behind the thousands of blocks of the INTERNAL kernel on the main stream
                # (round 1: the side stream's work ran in INTERNAL's tail, 104 us per iteration, profiles/r01l)
                self._streams.append(torch.cuda.Stream(self.device, priority=0 if not self._streams else -1))
                if len(self._streams) == 1:
                    torch.cuda.set_stream(self._streams[0])
            else:
                self._streams.append(None)

    def stream(self, idx: int = 0):
        self.setAvailableStreamSet(idx + 1)
        return self._streams[idx]

    def streamHandle(self, idx: int = 0) -> int:
        s = self.stream(idx)
        return 0 if s is None else s.cuda_stream

    def newEvent(self):
        return torch.cuda.Event(enable_timing=False) if self.runtime == Runtime.stream else None

    def sync(self, idx: int = 0) -> None:
        s = self.stream(idx)
        if s is not None:
            s.synchronize()

    def syncAll(self) -> None:
        for s in self._streams:
            if s is not None:
                s.synchronize()

    def barrier(self) -> None:
        if self.world > 1:
            dist.barrier(self.group)

    def toString(self) -> str:
        return f"Backend(rank {self.rank}/{self.world}, {self.device}, {self.runtime.value})"
