"""The static tile schedule of the tcgen05 kernels, checked on the CPU through the library's host-side view of the very
function the device code runs (`xtag_debug_tile_coords`, csrc/clip_tc.cu: tile_coords): every output tile is visited
exactly once for plain, n-slab and streamed (block-permuted) schedules; the CTAs of a cluster get consecutive m tiles
of ONE n tile (the precondition of the B-tile multicast); a streamed forward walks the column blocks in arrival order."""
import ctypes
import itertools

import pytest

from xtag_clip_b200 import _lib

BM, BN = 128, 256


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def coords(lib, M, N, slab, order, tile):
    m, n, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    arr = (ctypes.c_int * len(order))(*order) if order else None
    rc = lib.xtag_debug_tile_coords(M, N, slab, arr, tile, ctypes.byref(m), ctypes.byref(n), ctypes.byref(s))
    assert rc == 0
    return m.value, n.value, s.value


def tiles(M, N):
    return -(-M // BM), -(-N // BN)


@pytest.mark.parametrize("M,N", [(128, 256), (130, 300), (4096, 8192), (2049, 1025), (32768 // 8, 32768), (384, 256),
                                 (16 * 128 + 1, 5 * 256)])
@pytest.mark.parametrize("slab", [0, 1, 3, 16, 1000])
def test_every_tile_visited_once(lib, M, N, slab):
    num_m, num_n = tiles(M, N)
    seen = set()
    for t in range(num_m * num_n):
        m, n, _ = coords(lib, M, N, slab, None, t)
        assert 0 <= m < num_m and 0 <= n < num_n
        seen.add((m, n))
    assert len(seen) == num_m * num_n


def test_out_of_range_tile_is_rejected(lib):
    m, n = ctypes.c_int(), ctypes.c_int()
    assert lib.xtag_debug_tile_coords(256, 512, 0, None, 4, ctypes.byref(m), ctypes.byref(n), None) != 0
    assert lib.xtag_debug_tile_coords(256, 512, 0, None, -1, ctypes.byref(m), ctypes.byref(n), None) != 0


@pytest.mark.parametrize("M,N", [(256, 512), (512, 1024), (4096, 4096), (32768, 32768), (1024, 1000), (384, 768)])
@pytest.mark.parametrize("tune,want", [(0x4000, 2), (0x8000, 4), (0x0, 1)])
@pytest.mark.parametrize("slab", [0, 2])
def test_cluster_ctas_share_their_n_tile(lib, M, N, tune, want, slab):
    num_m, num_n = tiles(M, N)
    cl = lib.xtag_debug_pick_cluster(M, N, tune)
    assert cl in (1, 2, 4) and cl <= want and num_m % cl == 0
    if want > 1 and num_m % want == 0 and num_m * num_n >= 2 * want:
        assert cl == want
    for t0 in range(0, num_m * num_n, cl):
        group = [coords(lib, M, N, slab, None, t0 + i) for i in range(cl)]
        assert len({n for _, n, _ in group}) == 1                       # one shared B tile
        assert [m for m, _, _ in group] == list(range(group[0][0], group[0][0] + cl))


@pytest.mark.parametrize("W,b,M", [(8, 4096, 4096), (4, 8192, 8192), (2, 512, 512), (5, 256, 130)])
def test_streamed_forward_walks_blocks_in_arrival_order(lib, W, b, M):
    """slab = one column block of b columns; slab k of the schedule works on block order[k]."""
    N = W * b
    num_m, num_n = tiles(M, N)
    blk_tiles = b // BN
    for r in (0, W - 1, W // 2):
        order = [(r + j) % W for j in range(W)]
        seen, last_slab = set(), 0
        for t in range(num_m * num_n):
            m, n, s = coords(lib, M, N, blk_tiles, order, t)
            assert s >= last_slab                                        # slabs are visited in sequence
            last_slab = s
            assert n // blk_tiles == order[s]                            # ... and slab s is block order[s]
            seen.add((m, n))
        assert len(seen) == num_m * num_n
        # a CTA of a 148-wide grid meets the blocks in arrival order as well
        for cta in (0, 73, 147):
            blocks = [coords(lib, M, N, blk_tiles, order, t)[1] // blk_tiles for t in range(cta, num_m * num_n, 148)]
            dedup = [k for k, _ in itertools.groupby(blocks)]
            assert dedup == [blk for blk in order if blk in dedup] and len(set(dedup)) == len(dedup)


def work_item(lib, M, N, rows, slab, group, split, item):
    m, n, k = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.xtag_debug_work_item(M, N, rows, slab, group, split, item, ctypes.byref(m), ctypes.byref(n), ctypes.byref(k))
    assert rc == 0
    return m.value, n.value, k.value


@pytest.mark.parametrize("M,N", [(32768, 1024), (4096, 768), (330, 1320), (130, 300), (8192, 32768 // 8)])
@pytest.mark.parametrize("rows", [128, 256])
@pytest.mark.parametrize("group,split", [(0, 1), (1, 1), (1, 3), (0, 8), (5, 2)])
def test_pair_schedule_group_and_split_k_cover_every_work_item_once(lib, M, N, rows, group, split):
    """CTA-pair tiles (256 rows), the n-fastest order of the gradient GEMMs (group 1) and split-K work items: every
    (tile, K slice) exactly once; with group 1 the n tiles of one m block are consecutive; the K slices of a tile are
    adjacent work items (they run side by side)."""
    num_m, num_n = -(-M // rows), -(-N // BN)
    items = [work_item(lib, M, N, rows, 32, group, split, i) for i in range(num_m * num_n * split)]
    assert len(set(items)) == num_m * num_n * split
    assert all(0 <= m < num_m and 0 <= n < num_n and 0 <= k < split for m, n, k in items)
    for t in range(num_m * num_n):                                  # slices of one tile are consecutive
        sl = items[t * split:(t + 1) * split]
        assert len({(m, n) for m, n, _ in sl}) == 1 and [k for _, _, k in sl] == list(range(split))
    if group == 1 and num_n <= 32:
        tiles_only = items[::split]
        for mi in range(num_m):                                     # n-fastest: one m block, all its n tiles in a row
            run = tiles_only[mi * num_n:(mi + 1) * num_n]
            assert {m for m, _, _ in run} == {mi} and [n for _, n, _ in run] == list(range(num_n))
    assert lib.xtag_debug_work_item(M, N, 64, 0, 0, 1, 0, None, None, None) != 0
