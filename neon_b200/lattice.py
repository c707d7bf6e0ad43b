"""Lattice tables in the reference numbering.

D3Q19: benchmarks/lbm-lid-driven-cavity-flow/src/D3Q19.h:23-44 (c), :112-132 (t); opposite of q is q±10, rest = 9.
D3Q27: apps/lbmMultiRes/lattice.h:15-77 (rest = 0).
"""
from __future__ import annotations

import numpy as np


class D3Q19:
    Q = 19
    REST = 9
    c = np.array([[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-1, -1, 0], [-1, 1, 0], [-1, 0, -1], [-1, 0, 1], [0, -1, -1],
                  [0, -1, 1], [0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, -1, 0], [1, 0, 1], [1, 0, -1],
                  [0, 1, 1], [0, 1, -1]], np.int32)
    opp = np.array([q + 10 if q < 9 else (9 if q == 9 else q - 10) for q in range(19)], np.int32)
    t = np.array([1 / 18] * 3 + [1 / 36] * 6 + [1 / 3] + [1 / 18] * 3 + [1 / 36] * 6, np.float64)


class D3Q27:
    Q = 27
    REST = 0
    c = np.array([[0, 0, 0], [0, 0, -1], [0, 0, 1], [0, -1, 0], [0, -1, -1], [0, -1, 1], [0, 1, 0], [0, 1, -1],
                  [0, 1, 1], [-1, 0, 0], [-1, 0, -1], [-1, 0, 1], [-1, -1, 0], [-1, -1, -1], [-1, -1, 1], [-1, 1, 0],
                  [-1, 1, -1], [-1, 1, 1], [1, 0, 0], [1, 0, -1], [1, 0, 1], [1, -1, 0], [1, -1, -1], [1, -1, 1],
                  [1, 1, 0], [1, 1, -1], [1, 1, 1]], np.int32)
    opp = np.array([0, 2, 1, 6, 8, 7, 3, 5, 4, 18, 20, 19, 24, 26, 25, 21, 23, 22, 9, 11, 10, 15, 17, 16, 12, 14, 13],
                   np.int32)
    t = np.array([[8 / 27, 2 / 27, 1 / 54, 1 / 216][int(np.count_nonzero(v))] for v in c], np.float64)


def lattice(q: int):
    if q == 19:
        return D3Q19
    if q == 27:
        return D3Q27
    raise ValueError(f"unknown lattice D3Q{q}")


def crossing(q: int, direction: int):
    """Populations that cross a z face in ``direction`` (+1 up, -1 down): SURVEY.md §8e."""
    L = lattice(q)
    return [k for k in range(L.Q) if int(L.c[k, 2]) == direction]
