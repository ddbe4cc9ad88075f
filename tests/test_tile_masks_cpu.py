"""Host-only checks of the tile planning behind the symmetry-aware transforms (``qs_quarter_plan_tiles``; no device
is touched): a masked launch must visit EVERY tile that holds a wanted (row, column) pair -- dropping one would leave
holes in the result -- and should visit few tiles that hold none.  Brute force over rows and columns."""

import ctypes

import numpy as np
import pytest

torch = pytest.importorskip("torch")

F64, C128 = 0, 1


@pytest.fixture(scope="module")
def lib():
    from quantum_systems_b200 import _native
    from quantum_systems_b200.build import build

    build()
    return _native.load()


def plan(lib, X, K, W, a_dtype, m_dtype, x_inner, kind=0, strict=0, dh=1, mh=1, dl=1, ml=1, table=None):
    count = ctypes.c_int64(0)
    tptr = ctypes.c_void_p(table.ctypes.data) if table is not None else ctypes.c_void_p(0)
    args = (X, K, W, a_dtype, m_dtype, x_inner, kind, strict, dh, mh, dl, ml, tptr)
    assert lib.qs_quarter_plan_tiles(*args, ctypes.c_void_p(0), 0, ctypes.byref(count)) == 0
    tiles = np.zeros((count.value, 4), dtype=np.int64)
    assert lib.qs_quarter_plan_tiles(*args, ctypes.c_void_p(tiles.ctypes.data), count.value, ctypes.byref(count)) == 0
    return tiles


def covered(tiles, X, W):
    grid = np.zeros((X, W), dtype=bool)
    for x0, x1, w0, w1 in tiles:
        grid[x0 : x1 + 1, w0 : w1 + 1] = True
    return grid


@pytest.mark.parametrize("strict", [0, 1])
@pytest.mark.parametrize("dtypes", [(F64, F64), (C128, C128), (C128, F64), (F64, C128)])
@pytest.mark.parametrize("n,m", [(5, 7), (16, 16), (20, 33), (33, 20), (50, 64), (12, 130)])
def test_step2_mask_column_below_row_index(lib, n, m, dtypes, strict):
    """Step 2 of the single-GPU transform: rows (s, a, b), column r; wanted iff r < s (r <= s)."""
    X = m * n * n
    tiles = plan(lib, X, n, m, *dtypes, x_inner=n, kind=1, strict=strict, dl=n * n, ml=m)
    s = (np.arange(X) // (n * n))[:, None]
    r = np.arange(m)[None, :]
    wanted = (r < s) if strict else (r <= s)
    grid = covered(tiles, X, m)
    assert not (wanted & ~grid).any(), "a wanted (row, column) pair is in no launched tile"
    if m >= 128 and n * n >= 128:  # several column tiles: those right of every row index of a tile are skipped
        assert grid.mean() < 0.95


@pytest.mark.parametrize("strict", [0, 1])
@pytest.mark.parametrize("n,m", [(5, 7), (16, 16), (20, 33), (33, 20), (64, 50), (128, 40)])
def test_step3_mask_row_pairs(lib, n, m, strict):
    """Step 3: rows (r, s, a); wanted iff r < s (r <= s), every column."""
    X = m * m * n
    tiles = plan(lib, X, n, m, F64, F64, x_inner=n, kind=2, strict=strict, dh=m * n, mh=m, dl=n, ml=m)
    x = np.arange(X)
    r, s = x // (m * n), (x // n) % m
    wanted_rows = (r < s) if strict else (r <= s)
    grid = covered(tiles, X, m)
    assert grid[wanted_rows].all()
    if n >= 64:  # a 128-row tile then holds at most two (r, s) blocks: about half of the tiles are skipped
        assert grid.any(axis=1).mean() < 0.62


@pytest.mark.parametrize("world,rank", [(1, 0), (2, 1), (3, 0), (8, 5)])
@pytest.mark.parametrize("n,m", [(9, 9), (16, 24), (50, 50), (64, 48)])
def test_row_table_plan_of_the_sharded_step3(lib, n, m, world, rank):
    """Sharded step 3: rows (r_loc, s, a) with the cyclic pair rule in a row table; tiles without a kept row are
    not launched, every kept row is."""
    from quantum_systems_b200.sharded import block_partition, cyclic_wanted

    _, off = block_partition(m, world)
    R = off[rank + 1] - off[rank]
    if R == 0:
        pytest.skip("this rank owns no r")
    r = off[rank] + np.arange(R, dtype=np.int64)[:, None]
    s = np.arange(m, dtype=np.int64)[None, :]
    wanted = cyclic_wanted(r, s, m)
    table = np.full(R * m, -1, dtype=np.int64)
    table[wanted.reshape(-1)] = np.arange(int(wanted.sum())) * n
    X = R * m * n
    tiles = plan(lib, X, n, m, F64, F64, x_inner=n, table=table)
    grid = covered(tiles, X, m)
    kept_rows = np.repeat(table >= 0, n)
    assert grid[kept_rows].all()
    launched_rows = grid.any(axis=1)
    # every launched tile holds at least one kept row
    for x0, x1, _, _ in tiles:
        assert kept_rows[x0 : x1 + 1].any()
    if n >= 50:
        assert launched_rows.mean() < 0.75


@pytest.mark.parametrize("padded", [0, 1])
@pytest.mark.parametrize("dtypes", [(F64, F64), (C128, C128), (C128, F64)])
@pytest.mark.parametrize("n,m,planes", [(5, 7, 2), (16, 16, 4), (20, 66, 5), (40, 130, 3), (24, 200, 6), (9, 400, 1)])
def test_first_exchange_mask_cyclic_pairs(lib, n, m, planes, dtypes, padded):
    """First exchange of the sharded anti-symmetric schedule (mask kind 3): rows (s, a_loc, b), column r; wanted iff
    the cyclic pair rule computes (r, s) -- or, padded, the other member s ^ 1 of its aligned couple."""
    from quantum_systems_b200.sharded import cyclic_wanted

    X = m * planes * n
    tiles = plan(lib, X, n, m, *dtypes, x_inner=n, kind=3, strict=0 if padded else 1, dl=planes * n, ml=m)
    s = (np.arange(X) // (planes * n))[:, None]
    r = np.arange(m)[None, :]
    wanted = cyclic_wanted(r, s, m)
    if padded:
        wanted = wanted | cyclic_wanted(r, s ^ 1, m)
    grid = covered(tiles, X, m)
    assert not (wanted & ~grid).any(), "a wanted (row, column) pair is in no launched tile"
    if m >= 200 and planes * n >= 128:  # several column tiles: roughly half of them hold no wanted pair
        assert grid.mean() < 0.8
    for t in range(m):  # the library's rule is the schedule's rule
        assert lib.qs_cyclic_pair_wanted(3 % m, t, m) == int(cyclic_wanted(3 % m, t, m))
