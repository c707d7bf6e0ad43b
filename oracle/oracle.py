"""TEST INFRASTRUCTURE — numpy/ctypes front end of the CPU oracle (oracle/lbm_oracle.c).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  It is the checker, never the
product: ``neon_b200`` must not import anything from ``oracle/``.

All arrays are dense and unpadded: populations ``[q, z, y, x]``, classes and
wall masks ``[z, y, x]`` (the layout of oracle/ref_driver.cu dumps).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "liblbm_oracle.so")

BOUNCE, MOVING, BULK = 0, 1, 2
GEOM_CAVITY, GEOM_CAVITY_SPHERE, GEOM_FLOW_SPHERE = 0, 1, 2


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "lbm_oracle.c"), os.path.join(_HERE, "lbm_oracle_impl.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-s", "-f", "oracle/Makefile", "-B"], cwd=_ROOT)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def set_threads(n: int = 0) -> int:
    """Threads of the step loops (0 = all host cores).  The results do not depend on it."""
    n = n or (os.cpu_count() or 1)
    lib().olbm_set_threads(C.c_int(n))
    return n


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _sfx(dtype, compute=None) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32c64" if compute is not None and np.dtype(compute) == np.float64 else "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(dtype)


def tables(q: int):
    c = np.zeros((q, 3), np.int32)
    opp = np.zeros(q, np.int32)
    w = np.zeros(q, np.float64)
    lib().olbm_tables(C.c_int(q), _p(c), _p(opp), _p(w))
    return c, opp, w


def omega_cavity(n: int, re: float = 100.0, ulb: float = 0.04) -> float:
    """benchmarks/lbm-lid-driven-cavity-flow/src/Config.cpp:105-111"""
    nu = ulb * float(n - 2) / re
    return 1.0 / (3.0 * nu + 0.5)


def classify(geom: int, nx: int, ny: int, nz: int, sphere=None) -> np.ndarray:
    cls = np.empty((nz, ny, nx), np.int32)
    sp = None
    if sphere is not None:
        sp = np.asarray(sphere, np.float64)
    lib().olbm_classify(C.c_int(geom), C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(sp) if sp is not None else None, _p(cls))
    return cls


def wall_mask(q: int, cls: np.ndarray) -> np.ndarray:
    nz, ny, nx = cls.shape
    mask = np.empty((nz, ny, nx), np.uint32)
    f = lib().olbm_wall_mask
    f.restype = C.c_long
    bad = f(C.c_int(q), C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(np.ascontiguousarray(cls)), _p(mask))
    if bad:
        raise ValueError(f"{bad} bulk-cell neighbours fall outside the domain (geometry must be enclosed)")
    return mask


def init_pop(q: int, cls: np.ndarray, dtype, ulb: float = 0.04) -> np.ndarray:
    nz, ny, nx = cls.shape
    pop = np.empty((q, nz, ny, nx), dtype)
    getattr(lib(), "olbm_init_pop_" + _sfx(dtype))(
        C.c_int(q), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_double(ulb), _p(np.ascontiguousarray(cls)), _p(pop))
    return pop


def step(q: int, fin: np.ndarray, fout: np.ndarray, cls: np.ndarray, mask: np.ndarray, omega: float, compute=None):
    """One fused pull-stream + BGK iteration, fin -> fout (non-bulk cells of fout untouched)."""
    assert fin.dtype == fout.dtype and fin.flags.c_contiguous and fout.flags.c_contiguous
    assert cls.flags.c_contiguous and mask.flags.c_contiguous
    nz, ny, nx = cls.shape
    getattr(lib(), f"olbm_d3q{q}_step_" + _sfx(fin.dtype, compute))(
        C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(fin), _p(fout), _p(cls), _p(mask), C.c_double(omega))


def run(q: int, pop: np.ndarray, cls: np.ndarray, mask: np.ndarray, omega: float, iters: int, compute=None):
    """iters iterations with the two-field ping-pong of LbmIteration.h:56-61; returns the current input field."""
    a = pop.copy()
    b = pop.copy()
    for _ in range(iters):
        step(q, a, b, cls, mask, omega, compute)
        a, b = b, a
    return a


def rho_u(fin: np.ndarray, cls: np.ndarray, mask: np.ndarray, compute=None):
    nz, ny, nx = cls.shape
    rho = np.empty((nz, ny, nx), fin.dtype)
    u = np.empty((3, nz, ny, nx), fin.dtype)
    getattr(lib(), "olbm_d3q19_rho_u_" + _sfx(fin.dtype, compute))(
        C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(fin), _p(cls), _p(mask), _p(rho), _p(u))
    return rho, u


def read_ref_dump(path: str):
    """Parse a dump written by oracle/ref_driver.cu (the unmodified reference)."""
    with open(path, "rb") as f:
        hdr = np.frombuffer(f.read(32), np.int32)
        assert hdr[0] == 0x4E4C424D, "bad magic"
        nx, ny, nz, q, fpb, iters, geom = (int(v) for v in hdr[1:8])
        omega = float(np.frombuffer(f.read(8), np.float64)[0])
        dt = np.float32 if fpb == 4 else np.float64
        cells = nx * ny * nz
        pop = np.frombuffer(f.read(cells * q * fpb), dt).reshape(q, nz, ny, nx).copy()
        mask = np.frombuffer(f.read(cells * 4), np.uint32).reshape(nz, ny, nx).copy()
        cls = np.frombuffer(f.read(cells * 4), np.int32).reshape(nz, ny, nx).copy()
    return dict(nx=nx, ny=ny, nz=nz, q=q, iters=iters, geom=geom, omega=omega, pop=pop, mask=mask, cls=cls)
