"""How rx_front_kernel's tiles are dealt to CTAs (rx_make_deal / deal_lo / deal_owner, csrc/rx_kernels.cuh) -- host arithmetic,
walked through the C ABI's test aid: the ranges partition the tiles, the owner function is their inverse, and equal-length
channels of a batch are given whole CTAs (no range crosses a channel seam)."""
import ctypes as C

import numpy as np
import pytest

from gr_amps_b200 import capi


def deal(tiles, resident, nchan=1, equal_tiles=0):
    L = capi.lib()
    L.amps_b200_debug_deal.argtypes = [C.c_uint32] * 4 + [C.c_void_p] * 3
    grid = C.c_uint32(0)
    lo = np.zeros(1026, np.uint32)
    owner = np.zeros(tiles, np.uint32)
    capi.check(L.amps_b200_debug_deal(tiles, resident, nchan, equal_tiles, C.byref(grid), lo.ctypes.data_as(C.c_void_p),
                                      owner.ctypes.data_as(C.c_void_p)))
    return grid.value, lo[:grid.value + 1].astype(np.int64), owner.astype(np.int64)


def check_partition(tiles, grid, lo, owner):
    assert 1 <= grid <= 1024
    assert lo[0] == 0 and lo[-1] == tiles
    assert np.all(np.diff(lo) >= 0)
    want = np.repeat(np.arange(grid), np.diff(lo))
    assert np.array_equal(owner, want)


@pytest.mark.parametrize("resident", [1, 2, 7, 148, 296, 444, 1024, 5000])
def test_single_channel_ranges_partition_the_tiles(resident):
    rng = np.random.default_rng(resident)
    for tiles in [1, 2, 3, 4, 7, 14, 295, 296, 297, 437, 3496, 56320] + [int(v) for v in rng.integers(1, 200000, 12)]:
        grid, lo, owner = deal(tiles, resident)
        check_partition(tiles, grid, lo, owner)
        assert grid == min(resident, tiles, 1024)
        assert np.diff(lo).max() - np.diff(lo).min() <= 1            # even shares: nobody has two tiles more than anybody else


@pytest.mark.parametrize("nchan,tc", [(2, 1), (2, 437), (8, 437), (64, 4), (64, 7), (64, 55), (9, 30), (37, 8), (64, 1), (5, 100)])
def test_equal_channels_get_whole_ctas(nchan, tc):
    tiles = nchan * tc
    for resident in (296, 444, 148, 37, 1024):
        grid, lo, owner = deal(tiles, resident, nchan, tc)
        check_partition(tiles, grid, lo, owner)
        seams = np.arange(1, nchan) * tc
        crossing = [(a, b) for a, b in zip(lo[:-1], lo[1:]) if np.any((seams > a) & (seams < b))]
        p = min(resident, tiles, 1024) // nchan
        per_channel_mode = nchan <= min(resident, tiles, 1024) and 4 * nchan * min(p, tc) >= 3 * min(resident, tiles, 1024)
        if per_channel_mode:
            assert not crossing, "a CTA's range crosses a channel seam in per-channel mode"
            assert grid == nchan * min(p, tc)
            per = np.diff(lo).reshape(nchan, -1)
            assert np.all(per.sum(axis=1) == tc) and per.max() - per.min() <= 1
        else:
            assert grid == min(resident, tiles, 1024)


def test_unequal_channels_fall_back_to_one_range():
    grid, lo, owner = deal(1000, 296, 5, 0)
    check_partition(1000, grid, lo, owner)
    assert grid == 296
