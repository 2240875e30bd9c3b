"""Host logic of the paged KV cache through the C ABI (no kernels run, so no GPU needed): the end-index arithmetic
must equal the oracle's restatement of causal_model.py:277-300 bit for bit, the block table must reproduce the
reference's frame provenance (SURVEY §8c KATs), and valid pages must stay a physical prefix."""
import ctypes
import random

import pytest

from inferix_b200 import _lib
from oracle import wan_oracle as wo


class HostKV:
    """ifx_kv over fake (never dereferenced) buffers."""

    def __init__(self, num_pages, page_tokens, heads=2, head_dim=128):
        self.lib = _lib.load()
        self.h = ctypes.c_void_p()
        _lib.check(self.lib.ifx_kv_create(ctypes.byref(self.h), 0x10000, 0x20000, num_pages, page_tokens, heads, head_dim))
        self.num_pages = num_pages

    def plan(self, current_start, num_new, sink, windowed=True):
        p = _lib.KvPlan()
        _lib.check(self.lib.ifx_kv_plan_append(self.h, current_start, num_new, sink, int(windowed), ctypes.byref(p)))
        return p

    def state(self):
        g, l, n = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
        t = (ctypes.c_int32 * self.num_pages)()
        _lib.check(self.lib.ifx_kv_state(self.h, ctypes.byref(g), ctypes.byref(l), ctypes.byref(n), t, self.num_pages))
        return g.value, l.value, list(t[:n.value])

    def close(self):
        _lib.check(self.lib.ifx_kv_destroy(self.h))


def provenance(cache_frames, sink, nblocks, fs=4, block=3, forwards=2):
    kv = HostKV(cache_frames, fs)
    content = {}      # physical page -> source frame id
    snaps = []
    for b in range(nblocks):
        for _ in range(forwards):
            p = kv.plan(b * block * fs, block * fs, sink * fs)
            for i in range(p.num_pages):
                content[p.pages[i]] = b * block + i
        _, local_end, table = kv.state()
        assert sorted(table) == list(range(len(table))), "valid pages must be the physical prefix"
        assert local_end == len(table) * fs
        snaps.append([content[pg] for pg in table])
    kv.close()
    return snaps


def test_frame_provenance_kat():
    assert provenance(6, 0, 5) == [[0, 1, 2], [0, 1, 2, 3, 4, 5], [3, 4, 5, 6, 7, 8], [6, 7, 8, 9, 10, 11],
                                   [9, 10, 11, 12, 13, 14]]
    assert provenance(6, 1, 5) == [[0, 1, 2], [0, 1, 2, 3, 4, 5], [0, 4, 5, 6, 7, 8], [0, 7, 8, 9, 10, 11],
                                   [0, 10, 11, 12, 13, 14]]
    assert provenance(7, 1, 5) == [[0, 1, 2], [0, 1, 2, 3, 4, 5], [0, 3, 4, 5, 6, 7, 8], [0, 6, 7, 8, 9, 10, 11],
                                   [0, 9, 10, 11, 12, 13, 14]]


def test_index_kat():
    def run(cache_frames, sink, nblocks):
        kv = HostKV(cache_frames, 4)
        res = []
        for b in range(nblocks):
            first = None
            for _ in range(3):
                p = kv.plan(b * 12, 12, sink * 4)
                first = first or (p.local_start, p.local_end)
                assert (p.local_start, p.local_end) == first   # roll fires once per block
            g, l, _ = kv.state()
            res.append((g, l))
        kv.close()
        return res
    assert run(6, 0, 4) == [(12, 12), (24, 24), (36, 24), (48, 24)]
    assert run(6, 1, 4) == [(12, 12), (24, 24), (36, 24), (48, 24)]
    assert run(9, 0, 5) == [(12, 12), (24, 24), (36, 36), (48, 36), (60, 36)]


@pytest.mark.parametrize("seed", range(8))
def test_indices_match_oracle_randomised(seed):
    rnd = random.Random(seed)
    fs = rnd.choice([1, 4, 64, 1560, 3600])
    block = rnd.choice([1, 2, 3, 4])
    cache_frames = rnd.randint(block + 2, 30)
    sink = rnd.randint(0, 2)
    windowed = rnd.random() < 0.8
    kv = HostKV(cache_frames, fs)
    g = l = 0
    for b in range(40):
        for _ in range(rnd.randint(1, 4)):
            want = wo.plan_indices(cache_frames * fs, g, l, b * block * fs, block * fs, sink * fs, windowed)
            if want[0] < 0 or want[1] > cache_frames * fs:
                with pytest.raises(IndexError):
                    kv.plan(b * block * fs, block * fs, sink * fs, windowed)
                kv.close()
                return
            p = kv.plan(b * block * fs, block * fs, sink * fs, windowed)
            assert (p.local_start, p.local_end, p.global_end, p.num_evicted) == want
            g, l = want[2], want[1]
            assert kv.state()[:2] == (g, l)
    kv.close()


def test_long_horizon_no_growth():
    """BASELINE config 5 shape: 256 blocks through an 8-block window; table size and page set stay fixed."""
    fs, block, window = 3600, 3, 24
    kv = HostKV(window, fs)
    for b in range(256):
        for _ in range(2):
            p = kv.plan(b * block * fs, block * fs, 0)
        g, l, table = kv.state()
        assert g == (b + 1) * block * fs
        assert l == min((b + 1) * block, window) * fs
        assert len(table) == min((b + 1) * block, window) and sorted(table) == list(range(len(table)))
        assert p.num_evicted == 0                      # second forward of a block never evicts
    kv.close()


def test_unaligned_append_is_rejected():
    kv = HostKV(6, 4)
    with pytest.raises(NotImplementedError):
        kv.plan(0, 6, 0)
    with pytest.raises(NotImplementedError):
        kv.plan(2, 4, 0)
    kv.close()


def test_kv_map_identity_rows_and_refusals():
    """ifx_kv_map (MAGI caches, magi_kv_cache_manager.py:110-146: written in place, never evicted): maps logical
    tokens identity-wise, returns the row base pointers, refuses out-of-range requests and rotated tables."""
    kv = HostKV(64, 1)
    k, v = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(kv.lib.ifx_kv_map(kv.h, 40, ctypes.byref(k), ctypes.byref(v)))
    assert (k.value, v.value) == (0x10000, 0x20000)
    assert kv.state()[2] == list(range(40))
    _lib.check(kv.lib.ifx_kv_map(kv.h, 16, None, None))            # shrinking request: table keeps its 40 entries
    assert kv.state()[2] == list(range(40))
    _lib.check(kv.lib.ifx_kv_map(kv.h, 64, None, None))
    assert kv.state()[2] == list(range(64))
    with pytest.raises(IndexError):
        _lib.check(kv.lib.ifx_kv_map(kv.h, 65, None, None))
    kv.close()
    rolled = HostKV(4, 8)                                          # windowed cache: third block evicts -> rotated
    for b in range(3):
        rolled.plan(b * 16, 16, 0)
    with pytest.raises(NotImplementedError):
        _lib.check(rolled.lib.ifx_kv_map(rolled.h, 8, None, None))
    _lib.check(rolled.lib.ifx_kv_reset(rolled.h))                  # reset clears the rotation
    _lib.check(rolled.lib.ifx_kv_map(rolled.h, 8, None, None))
    rolled.close()
