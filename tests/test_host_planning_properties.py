"""Property tests (hypothesis) of the host planning entry points of the C ABI against the oracle: integer work, bit-exact.
The oracle itself is pinned to transformers' outputs by tests/test_oracle_golden.py."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from karanta_ocr_b200 import _lib, smart_resize
from oracle import preprocess_oracle as po
from oracle import vision_oracle as vo

dims = st.integers(min_value=1, max_value=6000)
budgets = st.sampled_from([(3136, 12845056), (3136, 1003520), (256 * 28 * 28, 1280 * 28 * 28), (56 * 56, 28 * 28 * 64), (784, 784 * 4)])


@settings(max_examples=600, deadline=None)
@given(h=dims, w=dims, budget=budgets)
def test_smart_resize_matches_oracle(h, w, budget):
    mn, mx = budget
    try:
        want = po.smart_resize(h, w, 28, mn, mx)
    except ValueError:
        with pytest.raises(ValueError):
            smart_resize(h, w, 28, mn, mx)
        return
    got = smart_resize(h, w, 28, mn, mx)
    assert got == tuple(want)
    assert got[0] % 28 == 0 and got[1] % 28 == 0
    n = _lib.load().kocr_num_patches(h, w, 14, 2, C.c_int64(mn), C.c_int64(mx))
    assert n == (got[0] // 14) * (got[1] // 14)


@settings(max_examples=120, deadline=None)
@given(in_size=st.integers(min_value=1, max_value=5000), out_units=st.integers(min_value=1, max_value=140),
       mode=st.sampled_from([po.RESIZE_PIL, po.RESIZE_ATEN]))
def test_filter_bank_matches_oracle(in_size, out_units, mode):
    out_size = 28 * out_units
    lib = _lib.load()
    k = lib.kocr_resample_ksize(in_size, out_size)
    assert k == po.resample_ksize(in_size, out_size)
    b = np.zeros((out_size, 2), dtype=np.int32)
    c = np.zeros((out_size, k), dtype=np.int32)
    prec = C.c_int()
    assert lib.kocr_resample_coeffs(in_size, out_size, mode, b.ctypes.data, c.ctypes.data, C.byref(prec)) == 0
    ob, oc, oprec = po.resample_coeffs(in_size, out_size, mode)
    assert prec.value == oprec and np.array_equal(b, ob) and np.array_equal(c, oc)
    assert (b[:, 0] >= 0).all() and (b[:, 0] + b[:, 1] <= in_size).all()  # every tap window lies inside the input


grids = st.lists(st.tuples(st.just(1), st.integers(1, 60).map(lambda v: 2 * v), st.integers(1, 60).map(lambda v: 2 * v)), min_size=1, max_size=5)


@settings(max_examples=80, deadline=None)
@given(grid=grids)
def test_index_tables_match_oracle(grid):
    lib = _lib.load()
    g = np.ascontiguousarray(np.asarray(grid, dtype=np.int64))
    total = int((g[:, 0] * g[:, 1] * g[:, 2]).sum())
    pos = np.zeros((total, 2), dtype=np.int32)
    assert lib.kocr_pos_ids(g.ctypes.data, len(g), 2, pos.ctypes.data) == 0
    assert np.array_equal(pos, vo.pos_ids(g).astype(np.int32))
    cu = np.zeros(len(g) + 1, dtype=np.int32)
    ncu = C.c_int()
    assert lib.kocr_cu_seqlens(g.ctypes.data, len(g), cu.ctypes.data, C.byref(ncu)) == 0
    assert np.array_equal(cu[:ncu.value], vo.cu_seqlens(g))
    wi = np.zeros(total // 4, dtype=np.int32)
    cuw = np.zeros(total // 4 + 1, dtype=np.int32)
    ncw = C.c_int()
    assert lib.kocr_window_index(g.ctypes.data, len(g), 112, 2, 14, wi.ctypes.data, cuw.ctypes.data, C.byref(ncw)) == 0
    owi, ocuw = vo.window_index(g, 112, 2, 14)
    assert np.array_equal(wi, owi.astype(np.int32)) and np.array_equal(cuw[:ncw.value], ocuw.astype(np.int32))
    assert sorted(wi.tolist()) == list(range(total // 4))  # a permutation of the 4-patch groups


@settings(max_examples=200, deadline=None)
@given(costs=st.lists(st.floats(min_value=0.5, max_value=1e6, allow_nan=False), min_size=0, max_size=200), world=st.integers(1, 8))
def test_shard_pages_is_a_balanced_partition(costs, world):
    from karanta_ocr_b200 import gather_pages, shard_pages
    shards = shard_pages(costs, world)
    assert len(shards) == world
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(len(costs)))                      # every page exactly once
    assert all(s == sorted(s) for s in shards)                  # each rank keeps input order
    if costs:
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) <= sum(costs) / world + max(costs) + 1e-6 * sum(costs)  # greedy LPT bound
    # the host-side gather puts per-rank results back in page order (single process: identity routing)
    merged = [None] * len(costs)
    for s in shards:
        part = gather_pages([f"page{i}" for i in s], s, len(costs))
        for i in s:
            merged[i] = part[i]
    assert merged == [f"page{i}" for i in range(len(costs))]
