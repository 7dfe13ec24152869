"""Host logic of the blocked sweep's work structure (cugraph_b200/csrc/graph_build.cu: plan_hot_units), called through
the host-only C entry cugraph_b200_debug_plan_hot_units — no GPU needed.  Checks the invariants the kernels rely on:

* sub-units tile the three slot spaces (full / half / quarter) and the seg_row space contiguously, in class order,
  one class per sub-unit, groups of 32 pieces, at most `unit_slots` slots per sub-unit;
* the fill records cover every piece exactly once, in sorted-piece order;
* units partition the sub-units, never mix blocks, are ordered by block, and are closed once they hold >= unit_slots
  slots (narrow slots count half);
* the per-CTA ranges are a monotone partition of the units, balanced to within one unit of the cost target.
"""
import ctypes as C

import numpy as np
import pytest

from cugraph_b200 import _capi


def _plan(counts, narrow, unit_slots=8192, sm_count=148, cold_cost=2.0):
    """counts: int array [(B+1), kinds] of pieces per class (the last block is the cold one)."""
    L = _capi.lib()
    nb, kinds = counts.shape
    cstart = np.zeros(nb * kinds + 1, dtype=np.int32)
    cstart[1:] = np.cumsum(counts.reshape(-1))
    cap = int(counts.sum() // 32 + counts.size + 16)
    totals = np.zeros(7, dtype=np.int64)
    subs = np.zeros((cap, 4), dtype=np.int32)
    fills = np.zeros((cap, 4), dtype=np.int32)
    units = np.zeros((cap, 4), dtype=np.int32)
    rng = np.zeros(sm_count + 1, dtype=np.int32)
    n_subs, n_units, err = C.c_size_t(), C.c_size_t(), C.c_void_p()
    f = L.cugraph_b200_debug_plan_hot_units
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p,
                  C.c_size_t, C.POINTER(C.c_void_p)]
    code = f(cstart.ctypes.data, nb - 1, int(narrow), unit_slots, sm_count, cold_cost, totals.ctypes.data,
             subs.ctypes.data, fills.ctypes.data, cap, C.byref(n_subs), units.ctypes.data, cap, C.byref(n_units),
             rng.ctypes.data, rng.size, C.byref(err))
    _capi.check(code, err, "cugraph_b200_debug_plan_hot_units")
    n_cta = int(totals[5])
    return cstart, totals, subs[:n_subs.value], fills[:n_subs.value], units[:n_units.value], rng[:n_cta + 1]


def _check(counts, narrow, unit_slots=8192, sm_count=148, cold_cost=2.0):
    nb, kinds = counts.shape
    B = nb - 1
    cstart, totals, subs, fills, units, rng = _plan(counts, narrow, unit_slots, sm_count, cold_cost)
    slots, rows, cold0, hslots, qslots, n_cta, sslots = [int(t) for t in totals]
    # ---- sub-units
    run = {"full": 0, "h": 0, "q": 0, "s": 0}
    row_run = 0
    piece = 0
    last_key = -1
    eq_slots = []          # unit-size equivalent of every sub-unit
    for (slot_begin, row_begin, n_groups, code), (p0, p1, blk, _) in zip(subs, fills):
        kind_space = {16: "h", 32: "q", 64: "s"}.get(int(code), "full")
        steps = 1 if code > 8 else code
        assert 1 <= steps <= 8 and n_groups >= 1
        assert slot_begin == run[kind_space] and row_begin == row_run
        assert p0 == piece and p0 < p1 <= p0 + 32 * n_groups and p1 > p0 + 32 * (n_groups - 1)
        # the pieces of a sub-unit belong to one class of one block, classes appear in key order
        if narrow:
            kind = {64: 0, 32: 1, 16: 2}.get(int(code), code + 2)
        else:
            kind = code - 1
        key = blk * kinds + kind
        assert key >= last_key and cstart[key] <= p0 and p1 <= cstart[key + 1]
        if blk == B:
            assert code <= 8          # the cold block never holds narrow pieces
        last_key = key
        n_slots = 32 * n_groups * steps
        assert n_slots <= max(unit_slots, 32 * steps) or code > 8
        if code > 8:
            assert 16 * n_groups <= max(unit_slots, 16)
        run[kind_space] += n_slots
        row_run += 32 * n_groups
        piece = p1
        eq_slots.append(16 * n_groups if code > 8 else n_slots)
        if kind_space == "full":  # full slots of the hot blocks come first, the cold block's (32-bit ids) after them
            assert slot_begin >= cold0 if blk == B else slot_begin + n_slots <= cold0
    assert piece == counts.sum()
    assert (run["full"], run["h"], run["q"], run["s"], row_run) == (slots, hslots, qslots, sslots, rows)
    if not narrow:
        assert hslots == 0 and qslots == 0 and sslots == 0
    # ---- units
    eq_slots = np.array(eq_slots, dtype=np.int64)
    nxt = 0
    last_blk = -1
    for i, (s0, s1, blk, _) in enumerate(units):
        assert s0 == nxt and s1 > s0
        assert (fills[s0:s1, 2] == blk).all() and blk >= last_blk
        size = int(eq_slots[s0:s1].sum())
        closes_block = (i + 1 == len(units)) or units[i + 1][2] != blk
        assert size >= unit_slots or closes_block          # only the last unit of a block may be short
        assert int(eq_slots[s0:s1 - 1].sum()) < unit_slots  # closed as soon as the threshold was reached
        nxt, last_blk = s1, blk
    assert nxt == len(subs)
    # ---- CTA ranges
    assert n_cta == max(1, min(sm_count, len(units)))
    assert rng[0] == 0 and rng[-1] == len(units) and (np.diff(rng) >= 0).all()
    if len(units):
        cost = np.array([eq_slots[s0:s1].sum() * (cold_cost if blk == B else 1.0) + 64.0 for s0, s1, blk, _ in units])
        csum = np.concatenate([[0.0], np.cumsum(cost)])
        for c in range(1, n_cta):
            target = csum[-1] * c / n_cta
            u = rng[c]
            assert csum[u] <= target + 1e-6 and (u == len(units) or csum[u + 1] > target - 1e-6)
    return totals, subs, units, rng


@pytest.mark.parametrize("narrow", [False, True])
def test_plan_random_histograms(narrow):
    r = np.random.default_rng(7 + int(narrow))
    kinds = 11 if narrow else 8
    for trial in range(40):
        nb = int(r.integers(1, 40))
        counts = r.integers(0, 3000, size=(nb, kinds)).astype(np.int64)
        counts[r.random(counts.shape) < 0.3] = 0
        if trial % 5 == 0:
            counts[0, kinds - 1] = int(r.integers(50_000, 400_000))   # a heavy class of full pieces in block 0
        if narrow:
            counts[nb - 1, :3] = 0   # the cold block never holds narrow pieces (graph_build.cu: k_hot_emit_pieces)
        if trial % 7 == 0:
            counts[nb - 1] = 0       # no cold block at all: every column block is hot
        _check(counts, narrow, unit_slots=int(r.choice([1024, 4096, 8192, 32768])), sm_count=int(r.choice([1, 8, 148])))


def test_plan_empty_and_tiny():
    for narrow in (False, True):
        kinds = 11 if narrow else 8
        totals, subs, units, rng = _check(np.zeros((3, kinds), dtype=np.int64), narrow)
        assert len(subs) == 0 and len(units) == 0 and list(rng) == [0, 0]
        one = np.zeros((2, kinds), dtype=np.int64)
        one[0, kinds - 8] = 1        # a single one-slot piece
        totals, subs, units, rng = _check(one, narrow)
        assert len(subs) == 1 and len(units) == 1 and int(totals[0]) == 32 and int(totals[1]) == 32


def test_plan_rmat24_like_shape():
    """Piece histogram of the shape RMAT-24 produces (181 hot blocks, block 0 heavy, sparse tail): the plan has a few
    thousand units, every CTA gets a contiguous range and walks at most a handful of blocks."""
    kinds, nb = 8, 182
    counts = np.zeros((nb, kinds), dtype=np.int64)
    counts[0] = [200_000, 150_000, 100_000, 80_000, 60_000, 50_000, 40_000, 1_500_000]
    for b in range(1, nb - 1):
        scale = 1.0 / (1 + b) ** 0.8
        counts[b] = (np.array([900_000, 250_000, 90_000, 40_000, 20_000, 10_000, 6_000, 60_000]) * scale).astype(np.int64)
    totals, subs, units, rng = _check(counts, False)
    assert 1000 < len(units) < 20000
    blocks_per_cta = [len(set(units[rng[c]:rng[c + 1], 2])) for c in range(len(rng) - 1)]
    assert max(blocks_per_cta) <= 40 and np.mean(blocks_per_cta) < 6
