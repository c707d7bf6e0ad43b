"""Host logic of the block-sparse grid (no GPU): blocks, connectivity, partitions, dense <-> block conversion.
Reference behaviour: libNeonDomain/include/Neon/domain/details/bGrid/{bGrid_imp.h:7-185, bPartition_imp.h:194-198}."""
import numpy as np
import pytest

import neon_b200 as nb
from neon_b200.bgrid import NO_BLOCK, B


@pytest.fixture(scope="module")
def bk():
    return nb.Backend(runtime=nb.Runtime.openmp)


def test_dense_block_roundtrip_and_padding(bk):
    g = nb.bGrid(bk, (20, 12, 17))  # not multiples of 8
    assert g.nb == (3, 2, 3) and g.n_blocks == 18 == g.n_blocks_alloc and g.n_down == g.n_up == 0
    f = g.newField("p", 3, np.float64)
    h = np.random.default_rng(1).random((3, 17, 12, 20))
    f.updateDeviceData(h)
    assert np.array_equal(f.updateHostData(), h)
    # element offset (q * n_alloc + blk) * 512 + z*64 + y*8 + x  (include/neon_lbm.h)
    blk = int(np.nonzero((g.block_coords == (1, 0, 2)).all(axis=1))[0][0])
    assert f.view3[2, blk, 3 * 64 + 5 * 8 + 1].item() == h[2, 8 + 3, 5, 16 + 1]


def test_connectivity_is_consistent(bk):
    act = lambda x, y, z: (x - 20) ** 2 + (y - 20) ** 2 + (z - 20) ** 2 > 15 ** 2  # a hole of inactive blocks in the middle
    g = nb.bGrid(bk, (40, 40, 40), active=act)
    assert g.n_blocks < 125
    info = g.info_host
    coords = g.block_coords
    for b in range(g.n_blocks):
        assert info[b, 13] == b  # (0,0,0) is the block itself
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    k = (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)  # bPartition_imp.h:194-198
                    n = info[b, k]
                    if n == NO_BLOCK:
                        continue
                    assert tuple(coords[n]) == (coords[b][0] + dz, coords[b][1] + dy, coords[b][2] + dx)
                    assert info[n, 26 - k] == b  # and back
        assert tuple(info[b, 27:30]) == (coords[b][2] * B, coords[b][1] * B, coords[b][0] * B)
    assert g.getNumActiveCells() == int(np.sum(act(*np.meshgrid(np.arange(40), np.arange(40), np.arange(40), indexing="ij"))))


def test_partitions_by_block_layers(bk):
    dim = (24, 16, 56)  # 7 block layers over 3 partitions: 3 + 2 + 2
    parts = [nb.bGrid(bk, dim, partition=(i, 3)) for i in range(3)]
    assert [p.layers for p in parts] == [(0, 3), (3, 5), (5, 7)]
    assert sum(p.n_blocks for p in parts) == 3 * 2 * 7
    for lo, hi in zip(parts[:-1], parts[1:]):
        # the i-th block of my highest layer is the i-th ghost-down block of the partition above, and vice versa
        up_mine = lo.block_coords[lo.n_blocks - lo.n_up:lo.n_blocks]
        ghost_theirs = hi.block_coords[hi.n_blocks:hi.n_blocks + hi.n_ghost_down]
        assert np.array_equal(up_mine, ghost_theirs)
        dn_theirs = hi.block_coords[:hi.n_down]
        ghost_mine = lo.block_coords[lo.n_blocks + lo.n_ghost_down:]
        assert np.array_equal(dn_theirs, ghost_mine)
    assert parts[0].n_down == 0 and parts[2].n_up == 0 and parts[1].n_down == parts[1].n_up == 6
    # a local block of the lowest layer sees its ghost block below through the connectivity
    p = parts[1]
    assert p.info_host[0, (0 + 1) + 3 * (0 + 1) + 9 * (-1 + 1)] == p.n_blocks
    with pytest.raises(ValueError):
        nb.bGrid(bk, (8, 8, 24), partition=(0, 3))  # one block layer per partition is not enough


def test_flag_classes_roundtrip(bk):
    g = nb.bGrid(bk, (20, 12, 17))
    fl = g.newFlagField()
    cls = np.random.default_rng(2).integers(0, 3, (17, 12, 20)).astype(np.int32)
    fl.setClasses(cls)
    assert np.array_equal(fl.classes(), cls)
    assert not fl.masks().any()
