"""Host-side slack-fill scheduler of the attention kernel (ifx_attention_plan_info): every (item, key tile) is covered
exactly once, partial slots are dense and consecutive per item, CTAs stay within the segment limit, a CTA's own piece
comes first, and the plan is produced exactly for the single-wave shard shapes.  No GPU needed (host code)."""
import ctypes as C

import pytest

from inferix_b200 import _lib


def plan(q_rows, heads, n_tiles, n_old, sms=148):
    lib = _lib.load()
    grid, mk, bm, mean = C.c_int32(), C.c_double(), C.c_double(), C.c_double()
    cap = 8192
    seg = (C.c_int32 * (7 * cap))()
    _lib.check(lib.ifx_attention_plan_info(q_rows, heads, n_tiles, n_old, sms, C.byref(grid), C.byref(mk), C.byref(bm),
                                           C.byref(mean), seg, cap))
    rows = []
    if grid.value:
        for i in range(cap):
            if seg[7 * i] == -1:
                break
            rows.append(tuple(seg[7 * i:7 * i + 7]))
    return grid.value, mk.value, bm.value, mean.value, rows


# (rows per rank, heads, key tiles, resident tiles): 720p shards at 8 / 4 ranks (steady state, filling window), 480p at 8
@pytest.mark.parametrize("q_rows,heads,tiles,old", [(1350, 12, 677, 592), (1350, 12, 675, 675), (2700, 12, 677, 592),
                                                    (1350, 12, 255, 170), (585, 12, 258, 221), (2700, 12, 675, 675)])
def test_slack_fill_covers_every_tile_once(q_rows, heads, tiles, old):
    grid, makespan, base, mean, rows = plan(q_rows, heads, tiles, old)
    assert 0 < grid <= 148 and makespan < 0.97 * base and makespan <= 1.15 * mean
    pairs = (q_rows + 255) // 256
    seen, per_cta, per_item = set(), {}, {}
    for cta, item, a, n, c, m, slot in rows:
        assert 0 <= item < pairs * heads and n > 0 and m >= 0
        per_cta.setdefault(cta, []).append((item, a, n, c, m, slot))
        for t in list(range(a, a + n)) + list(range(c, c + m)):
            assert 0 <= t < tiles and (item, t) not in seen
            seen.add((item, t))
        per_item.setdefault(item, []).append(slot)
    assert len(seen) == pairs * heads * tiles
    assert max(len(v) for v in per_cta.values()) <= 12
    slots = sorted(s for v in per_item.values() for s in v if s >= 0)
    assert slots == list(range(len(slots)))                       # dense numbering
    for item, ss in per_item.items():
        if len(ss) == 1:
            assert ss == [-1]                                     # an uncut item writes the output directly
        else:
            assert sorted(ss) == list(range(min(ss), min(ss) + len(ss)))
    # a CTA's first segment is its own (long) piece; the shed chunks that follow are short
    for segs in per_cta.values():
        lens = [n + m for (_i, _a, n, _c, m, _s) in segs]
        assert all(x <= lens[0] for x in lens[1:]) or len(segs) == 1 or lens[0] >= 8


def test_slack_fill_only_for_single_wave_shapes():
    for q_rows, expect in ((1350, True), (2700, True), (5400, False), (10800, False)):
        grid, *_ = plan(q_rows, 12, 677, 592)
        assert (grid > 0) == expect, q_rows
