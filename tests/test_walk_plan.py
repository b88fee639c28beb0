"""The host-side decomposition of a walker launch (heatWalkKernel: b200_heat2d_walk_plan_query, no device needed): for any
field geometry the front segments and the interior segments must cover every output row exactly once, the column windows
every column this rank owns, and the walker count must be what the kernel's index arithmetic expects."""
import ctypes as C

import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from alpaka_b200 import _lib

TOP, BOTTOM, LEFT, RIGHT = 1, 2, 4, 8


def query(ny, nx, pad_y, pad_x, edges, levels, slots):
    plan = _lib.Heat2dWalkPlan()
    rc = _lib.load().b200_heat2d_walk_plan_query(ny, nx, pad_y, pad_x, edges, levels, slots, C.byref(plan))
    assert rc == 0
    return plan


def check(plan, ny, nx, pad_y, pad_x, edges, levels):
    lo_y, hi_y = pad_y, ny + pad_y - 1
    out_lo = lo_y - (1 if edges & TOP else 0)  # ring row on a physical side
    out_hi = hi_y + (1 if edges & BOTTOM else 0)
    covered = {}
    for k in range(plan.n_front):
        assert plan.front_y0[k] < plan.front_y1[k]
        for j in range(plan.front_y0[k], plan.front_y1[k]):
            covered[j] = covered.get(j, 0) + 1
    segs = plan.n_segments - plan.n_front
    assert plan.segment_rows >= 1
    for s_ in range(segs):
        y0 = plan.interior_y0 + s_ * plan.segment_rows
        y1 = min(y0 + plan.segment_rows, plan.interior_y1)
        assert y0 < y1, "an empty interior segment"
        for j in range(y0, y1):
            covered[j] = covered.get(j, 0) + 1
    assert sorted(covered) == list(range(out_lo, out_hi + 1)), "output rows not covered"
    assert set(covered.values()) == {1}, "an output row belongs to two segments"
    # strips: exactly the sides with a neighbour, pad_y rows each, flagged
    n_strip = bin(plan.front_is_strip).count("1")
    assert n_strip == (0 if edges & TOP else 1) + (0 if edges & BOTTOM else 1)
    for k in range(plan.n_front):
        if (plan.front_is_strip >> k) & 1:
            assert plan.front_y1[k] - plan.front_y0[k] == pad_y
    # columns: window w stores [w * W, (w + 1) * W); the owned columns run from lo_x - ring to hi_x + ring
    W = plan.window_columns
    assert W == 128 - 8 * ((levels + 3) // 4)
    hi_x = nx + pad_x - 1
    assert plan.n_windows * W >= hi_x + 1 + (1 if edges & RIGHT else 0)
    assert (plan.n_windows - 1) * W <= hi_x + 1, "a window that stores nothing this rank owns"
    assert 1 + plan.n_edge_right <= plan.n_windows
    if not plan.split:
        assert plan.n_walkers == plan.n_segments * plan.n_windows
    else:
        n_edge = 1 + plan.n_edge_right
        assert plan.n_walkers == n_edge * plan.n_segments + (plan.n_windows - n_edge) * plan.n_front


@settings(max_examples=400, deadline=None)
@given(ny=st.integers(1, 3000), nx=st.integers(1, 3000), levels=st.sampled_from([4, 6, 8]), slots=st.sampled_from([8, 96, 1184, 1776, 2368]))
def test_stand_alone_fields(ny, nx, levels, slots):
    check(query(ny, nx, 1, 1, TOP | BOTTOM | LEFT | RIGHT, levels, slots), ny, nx, 1, 1, 15, levels)


@settings(max_examples=400, deadline=None)
@given(k=st.integers(2, 400), nx=st.integers(1, 3000), levels=st.sampled_from([4, 6, 8]), extra=st.integers(0, 2), top=st.booleans(), bottom=st.booleans(),
       slots=st.sampled_from([96, 1776]))
def test_row_slabs(k, nx, levels, extra, top, bottom, slots):
    pad = levels + 2 * (extra if levels + 2 * extra <= 8 else 0)  # ghost rows at least as deep as the launch
    ny = pad * k
    edges = LEFT | RIGHT | (TOP if top else 0) | (BOTTOM if bottom else 0)
    check(query(ny, nx, pad, 1, edges, levels, slots), ny, nx, pad, 1, edges, levels)


@settings(max_examples=400, deadline=None)
@given(ky=st.integers(2, 300), kx=st.integers(2, 300), levels=st.sampled_from([4, 6, 8]), edges=st.integers(0, 15), slots=st.sampled_from([96, 1776]))
def test_deep_tiles(ky, kx, levels, edges, slots):
    pad = levels
    ny, nx = pad * ky, pad * kx
    check(query(ny, nx, pad, pad, edges, levels, slots), ny, nx, pad, pad, edges, levels)


def test_known_answers_and_refusals():
    # 16384^2 stand-alone, four levels, 3 CTAs x 4 walkers x 148 SMs: 137 windows of 120 columns, window 0 and the last one edge windows
    p = query(16384, 16384, 1, 1, 15, 4, 1776)
    assert (p.window_columns, p.n_windows, p.n_edge_right, p.n_front) == (120, 137, 1, 0)
    assert (p.interior_y0, p.interior_y1) == (0, 16386)
    waves = p.n_walkers / 1776
    assert 3 <= waves <= 9, waves  # a few waves, not exactly one
    plan = _lib.Heat2dWalkPlan()
    lib = _lib.load()
    assert lib.b200_heat2d_walk_plan_query(0, 64, 1, 1, 15, 4, 96, C.byref(plan)) == -1
    assert lib.b200_heat2d_walk_plan_query(64, 64, 1, 1, 15, 3, 96, C.byref(plan)) == -1
    assert lib.b200_heat2d_walk_plan_query(64, 64, 1, 1, 15, 4, 96, None) == -1
