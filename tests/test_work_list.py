"""Host-side work lists of the persistent tensor-core products (no GPU needed): every (tile, split slice) of a product is
assigned to exactly one CTA pair, tiles above the diagonal of a lower-triangular output are skipped, the k-block counts
follow the triangle trimming of the kernels, and the pairs are balanced."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gp-derivatives-variational-inference_b200"))
from dsvgp_b200 import _lib  # noqa: E402


def work_list(M, N, K, a_tri, c_lower, nsplit, pairs, bke):
    fn = _lib._lib.dsvgp_tc_work_list          # prototype parsed from include/dsvgp_b200.h
    size = fn(M, N, K, a_tri, c_lower, nsplit, pairs, bke, None, 0)
    assert size > 0
    buf = np.zeros(size, dtype=np.int32)
    assert fn(M, N, K, a_tri, c_lower, nsplit, pairs, bke, ctypes.c_void_p(buf.ctypes.data), size) == size
    n_off = (pairs + 1 + 3) & ~3
    offs = buf[: pairs + 1]
    items = buf[n_off:].reshape(-1, 4)
    return offs, items


def tile_range(M, K, a_tri, bke, mt):
    nkb = -(-K // bke)
    kb0, kb1 = 0, nkb
    if a_tri == 1:
        kb1 = min(nkb, (mt * 256 + 256 + bke - 1) // bke)
    if a_tri == 2:
        kb0 = mt * 256 // bke
    return kb0, max(kb1, kb0)


CASES = [
    # M, N, K, a_tri, c_lower, nsplit, pairs, bke
    (3072, 49152, 3072, 1, 0, 1, 74, 64),     # A = W K_zx at C3
    (3072, 49152, 3072, 0, 0, 1, 74, 64),     # C = D A
    (3072, 49152, 3072, 2, 0, 1, 74, 64),     # dK_zx = W^T dA
    (3072, 3072, 49152, 0, 1, 8, 74, 64),     # Gram product, split-K
    (3200, 3200, 6144, 0, 1, 4, 74, 64),      # C4 Gram product: ragged last row tile
    (3200, 8192, 3200, 1, 0, 1, 74, 64),      # C4: ragged last row tile
    (3072, 3072, 3072, 2, 0, 1, 74, 32),      # 3xTF32 tail product
    (300, 1000, 300, 1, 0, 1, 74, 64),        # fewer items than pairs
    (512, 512, 1536, 0, 1, 2, 74, 64),        # split-K with more pairs than a tile has slices: uniform slices
    (3072, 49152, 3072, 1, 0, 1, 7, 64),      # a small part
]


@pytest.mark.parametrize("M,N,K,a_tri,c_lower,nsplit,pairs,bke", CASES)
def test_work_list_covers_and_balances(M, N, K, a_tri, c_lower, nsplit, pairs, bke):
    offs, items = work_list(M, N, K, a_tri, c_lower, nsplit, pairs, bke)
    T, NT = -(-M // 256), -(-N // 256)
    tiles = {(mt, nt) for mt in range(T) for nt in range(NT) if not (c_lower and nt > mt)}
    assert offs[0] == 0 and offs[pairs] == len(items) and np.all(np.diff(offs) >= 0)
    kb0 = items[:, 3] & 0xFFFF
    kb1 = (items[:, 3].astype(np.int64) >> 16) & 0xFFFF
    assert np.all(kb1 >= kb0)
    pieces = {}
    for (mt, nt, z, _), a, b in zip(items.tolist(), kb0.tolist(), kb1.tolist()):
        assert (mt, nt) in tiles and 0 <= z < nsplit
        assert z not in pieces.setdefault((mt, nt), {})          # a (tile, slice) is written exactly once
        pieces[(mt, nt)][z] = (a, b)
    assert set(pieces) == tiles
    nslices = {len(v) for v in pieces.values()}
    assert len(nslices) == 1                                     # every tile writes the same slices (the reduction sums all)
    ns = nslices.pop()
    assert ns <= nsplit
    for (mt, nt), segs in pieces.items():
        assert sorted(segs) == list(range(ns))
        lo, hi = tile_range(M, K, a_tri, bke, mt)
        cover = sorted((a, b) for a, b in segs.values() if b > a)
        pos = lo
        for a, b in cover:                                       # the non-empty pieces tile the k range of the tile
            assert a == pos
            pos = b
        assert pos == hi
    loads = np.array([(kb1 - kb0)[offs[q]: offs[q + 1]].sum() + 3 * (offs[q + 1] - offs[q]) for q in range(pairs)])
    longest = int((kb1 - kb0).max())
    if nsplit > 1 and ns < nsplit or len(items) >= 4 * pairs:
        if nsplit > 1 and ns <= 2:                               # stream-K: equal runs of k-blocks
            assert loads.max() - loads.min() <= 4 * 3 + 2, (loads.min(), loads.max())
        elif len(set((kb1 - kb0).tolist())) > 2:                 # longest-first tail: spread of about one light item
            assert loads.max() - loads.min() <= longest // 2 + 3, (loads.min(), loads.max())
        else:                                                    # equal items: at most one item of difference
            assert loads.max() - loads.min() <= longest + 3
