"""Problem set-up: lid-driven cavity and sphere geometries, initial populations.

Host path (mirror of RunCavityTwoPop.cu:159-242: forEachActiveCell on host arrays, then updateDeviceData):
``host_classes`` / ``host_populations`` build the global arrays with numpy.
Device path: ``FlagField.classify`` + ``computeWallNghMask`` + ``init_populations`` never touch the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _capi as capi
from .dgrid import FlagField, dField, dGrid
from .lattice import lattice

CAVITY, CAVITY_SPHERE, FLOW_SPHERE = capi.GEOM_CAVITY, capi.GEOM_CAVITY_SPHERE, capi.GEOM_FLOW_SPHERE


def default_sphere(dim: Sequence[int]):
    nx, ny, nz = dim
    return (0.45 * nx, 0.55 * ny, 0.5 * nz, min(nx, ny, nz) / 5.0)


def host_classes(geom: int, dim: Sequence[int], sphere: Optional[Sequence[float]] = None) -> np.ndarray:
    """Cell classes [nz, ny, nx] (RunCavityTwoPop.cu:208-224; sphere cases: SURVEY.md §8d from
    apps/lbmMultiRes/flowOverShape.h:64-100,165-175)."""
    nx, ny, nz = dim
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij", sparse=True)
    cx, cy, cz, R = sphere if sphere is not None else default_sphere(dim)
    dx, dy, dz = x - cx, y - cy, z - cz
    in_sphere = (dx * dx + dy * dy) + dz * dz < R * R
    edge = (x == 0) | (x == nx - 1) | (y == 0) | (y == ny - 1) | (z == 0) | (z == nz - 1)
    cls = np.full((nz, ny, nx), capi.BULK, np.int32)
    if geom in (CAVITY, CAVITY_SPHERE):
        if geom == CAVITY_SPHERE:
            cls[np.broadcast_to(in_sphere, cls.shape)] = capi.BOUNCE_BACK
        cls[np.broadcast_to(edge, cls.shape)] = capi.BOUNCE_BACK
        cls[:, ny - 1, :] = capi.MOVING_WALL
    elif geom == FLOW_SPHERE:
        cls[:, :, 0] = capi.MOVING_WALL
        cls[np.broadcast_to(in_sphere, cls.shape)] = capi.BOUNCE_BACK
        walls = (y == 0) | (y == ny - 1) | (z == 0) | (z == nz - 1) | (x == nx - 1)
        cls[np.broadcast_to(walls, cls.shape)] = capi.BOUNCE_BACK
    else:
        raise ValueError(geom)
    return cls


def host_populations(q: int, cls: np.ndarray, dtype, ulb: float = 0.04) -> np.ndarray:
    """Initial populations [q, nz, ny, nx]: bulk t_k, bounceBack 0, movingWall -6 t_k ulb (c_k . (1,0,0))
    (RunCavityTwoPop.cu:168-206; D3Q27: apps/lbmMultiRes/lidDrivenCavity.h:56-76)."""
    L = lattice(q)
    dtype = np.dtype(dtype)
    pop = np.empty((q,) + cls.shape, dtype)
    bulk, moving = cls == capi.BULK, cls == capi.MOVING_WALL
    for k in range(q):
        if q == 19:
            wall = dtype.type(-6.0 * L.t[k] * ulb * (L.c[k, 0] * 1.0 + L.c[k, 1] * 0.0 + L.c[k, 2] * 0.0))
        else:
            v = dtype.type(0)
            for d, uw in enumerate((ulb, 0.0, 0.0)):
                v = dtype.type(float(v) + float(L.c[k, d]) * uw)
            wall = dtype.type(float(v) * (-6.0 * L.t[k]))
        pop[k] = np.where(bulk, dtype.type(L.t[k]), np.where(moving, wall, dtype.type(0)))
    return pop


def init_populations(field: dField, flag: FlagField, q: int, ulb: float = 0.04, stream_idx: int = 0) -> None:
    """Device-side initial populations of ``field`` (all planes, ghosts included) from the classes in ``flag``."""
    g = field.grid
    d = g.desc(None, field, flag)
    sym = f"nlbm_{g.kind}_init_pop_f32" if field.dtype == np.float32 else f"nlbm_{g.kind}_init_pop_f64"
    capi.call(sym, C.byref(d), q, ulb, g.backend.streamHandle(stream_idx))
    if hasattr(field, "commitWalls"):
        field.commitWalls(stream_idx)


def setup_device(grid, q: int, dtype, geom: int = CAVITY, sphere=None, ulb: float = 0.04):
    """Everything on the device: (pop0, pop1, flag) ready to iterate."""
    pop0 = grid.newField("pop0", q, dtype)
    pop1 = grid.newField("pop1", q, dtype)
    flag = grid.newFlagField("flag", like=pop0)
    flag.classify(geom, sphere)
    flag.computeWallNghMask(q)
    init_populations(pop0, flag, q, ulb)
    init_populations(pop1, flag, q, ulb)
    return pop0, pop1, flag


def setup_host(grid, q: int, dtype, cls: np.ndarray, pop: np.ndarray):
    """The reference's flow: host arrays -> updateDeviceData -> wall mask on the device (RunCavityTwoPop.cu:226-241)."""
    pop0 = grid.newField("pop0", q, dtype)
    pop1 = grid.newField("pop1", q, dtype)
    flag = grid.newFlagField("flag", like=pop0)
    flag.setClasses(cls)
    pop0.updateDeviceData(pop)
    pop1.updateDeviceData(pop)
    flag.computeWallNghMask(q)
    return pop0, pop1, flag
