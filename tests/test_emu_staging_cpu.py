"""Graph staging on the CPU: the real staging code (capi_graph.cu, graph_build.cu — renumbering, the packed-key sort,
binning, the piece layout of the blocked sweep, the experimental narrow and ELL layouts) compiled as plain C++ against
the host emulation shim in emu/ and driven through the real C ABI with numpy arrays.  The staging kernels are
data-parallel loops without intra-block communication, so executing every "thread" of a launch in turn is exact.

Checked against numpy: the stored graph is the input multigraph (external ids), rows are degree-descending with sorted
neighbours and correct segment bounds; the piece layout reproduces every (row, source[, weight]) of the degree >= 32
rows exactly once, padding only where allowed; the ELL copy + k_spmv_low_ell reproduce the SpMV of the degree < 32 rows.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INT32, INT64, FLOAT32, FLOAT64 = 2, 3, 8, 9


@pytest.fixture(scope="module")
def emu():
    sys.path.insert(0, os.path.join(ROOT, "emu"))
    import build_emu
    try:
        path = build_emu.build()
    except Exception as e:  # no host compiler: nothing to emulate with
        pytest.skip(f"emulation build unavailable: {e}")
    L = C.CDLL(path)
    L.cugraph_create_resource_handle.restype = C.c_void_p
    L.cugraph_create_resource_handle.argtypes = [C.c_void_p]
    L.cugraph_type_erased_device_array_view_create.restype = C.c_void_p
    L.cugraph_type_erased_device_array_view_create.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    L.cugraph_type_erased_device_array_view_free.argtypes = [C.c_void_p]
    L.cugraph_error_message.restype = C.c_char_p
    L.cugraph_error_message.argtypes = [C.c_void_p]
    L.cugraph_graph_free.argtypes = [C.c_void_p]
    L.emu_graph_primary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.emu_sweep_layout.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.emu_reload_tuning.argtypes = [C.c_void_p]
    L.handle = L.cugraph_create_resource_handle(None)
    assert L.handle
    return L


class Props(C.Structure):
    _fields_ = [("is_symmetric", C.c_int), ("is_multigraph", C.c_int)]


def make_edges(V, E, seed, weighted=False, id_offset=0):
    """power-law-ish destinations AND sources (hubs on both sides), multi-edges and self-loops included"""
    r = np.random.default_rng(seed)
    dst = np.minimum((V * r.random(E) ** 3.0).astype(np.int64), V - 1)
    src = np.minimum((V * r.random(E) ** 2.0).astype(np.int64), V - 1)
    perm = r.permutation(V)                     # external ids carry no degree information
    src, dst = perm[src] + id_offset, perm[dst] + id_offset
    w = (r.random(E).astype(np.float32) + 0.25) if weighted else None
    return src.astype(np.int32), dst.astype(np.int32), w


def create_graph(L, src, dst, w, **flags):
    L.emu_reload_tuning(C.c_void_p(L.handle))   # the knobs are read from the environment per handle; tests change it per case
    views = [L.cugraph_type_erased_device_array_view_create(a.ctypes.data, a.size, t) if a is not None else None
             for a, t in ((src, INT32), (dst, INT32), (w, FLOAT32))]
    g, err = C.c_void_p(), C.c_void_p()
    code = L.cugraph_graph_create_with_times_sg(
        C.c_void_p(L.handle), C.byref(Props(0, 1)), None, C.c_void_p(views[0]), C.c_void_p(views[1]),
        C.c_void_p(views[2]) if views[2] else None, None, None, None, None,
        1, 1, int(flags.get("drop_self_loops", 0)), int(flags.get("drop_multi_edges", 0)), int(flags.get("symmetrize", 0)), 0,
        C.byref(g), C.byref(err))
    assert code == 0, L.cugraph_error_message(err)
    for v in views:
        if v:
            L.cugraph_type_erased_device_array_view_free(v)
    return g


def as_np(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(int(n),))


def primary(L, g):
    ints = (C.c_int64 * 8)()
    seg = (C.c_int32 * 8)()
    ptrs = (C.c_void_p * 5)()
    assert L.emu_graph_primary(g, ints, seg, ptrs) == 0
    n_rows, nnz, offs64, nnz_hi, nv, weighted, wsize = [int(x) for x in ints[:7]]
    assert not offs64
    off = as_np(ptrs[0], n_rows + 1, np.int32)
    idx = as_np(ptrs[1], nnz, np.int32)
    w = as_np(ptrs[2], nnz, np.float32) if weighted else None
    ext = as_np(ptrs[3], nv, np.int32)
    return dict(n_rows=n_rows, nnz=nnz, nnz_hi=nnz_hi, nv=nv, off=off, idx=idx, w=w, ext=ext, seg=list(seg))


def check_csr(P, src, dst, w):
    off, idx, ext = P["off"], P["idx"], P["ext"]
    deg = np.diff(off)
    assert (deg[:-1] >= deg[1:]).all()                                  # degree-descending rows = the binning
    for k, t in enumerate((32, 16, 8, 4, 2, 1, 0)):
        assert P["seg"][k] == int((deg >= t).sum())
    assert P["nnz_hi"] == int(off[P["seg"][0]])
    rows = np.repeat(np.arange(P["n_rows"]), deg)
    same_row = rows[1:] == rows[:-1]
    assert (idx[1:][same_row] >= idx[:-1][same_row]).all()               # neighbours ascending inside a row
    # the stored graph is the input multigraph (rows = destinations, entries = sources), on external ids
    got = np.stack([ext[rows], ext[idx]], 1)
    exp = np.stack([dst, src], 1)
    if w is None:
        key = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
        assert (key(got) == key(exp)).all()
    else:
        o1 = np.lexsort((P["w"], got[:, 1], got[:, 0]))
        o2 = np.lexsort((w, exp[:, 1], exp[:, 0]))
        assert (got[o1] == exp[o2]).all() and (P["w"][o1] == w[o2]).all()
    assert sorted(set(ext.tolist())) == sorted(set(src.tolist()) | set(dst.tolist()))


def lds_wavefronts(ids):
    """shared-memory wavefronts of one warp-wide 4-byte gather: max over the 32 banks of the distinct addresses on a bank"""
    u = np.unique(ids)
    return int(np.bincount(u & 31, minlength=32).max())


def hot_pieces(L, g, P, bank_order=False, stats=None):
    """(row, col[, w]) triples reconstructed from the piece layout + structural checks.  bank_order: the experimental
    layout whose slots are ordered by shared-memory bank (padding anywhere in a slot, on any of the 64 zero columns).
    stats: dict that receives the LDS count and the wavefront count of the 16-bit full-slot classes."""
    ints = (C.c_int64 * 12)()
    ptrs = (C.c_void_p * 10)()
    rc = L.emu_hot_layout(C.c_void_p(L.handle), g, ints, ptrs)
    assert rc == 0, f"emu_hot_layout returned {rc}"
    W, B, n_hi, nnz_hi, n_hot, n_slots, n_subs, n_units, n_cta, narrow, es = [int(x) for x in ints[:11]]
    assert n_hi == P["seg"][0] and nnz_hi == P["nnz_hi"]
    subs = as_np(ptrs[4], 4 * n_subs, np.int32).reshape(-1, 4)
    units = as_np(ptrs[5], 4 * n_units, np.int32).reshape(-1, 4)
    rng = as_np(ptrs[6], n_cta + 1, np.int32)
    assert rng[0] == 0 and rng[-1] == n_units and (np.diff(rng) >= 0).all()
    n_rows_seg = int(subs[:, 2].sum()) * 32
    seg_row = as_np(ptrs[3], n_rows_seg, np.int32)
    idx16 = as_np(ptrs[0], n_hot * 8, np.uint16)
    idx32 = as_np(ptrs[1], (n_slots - n_hot) * 8, np.int32)
    sw = as_np(ptrs[2], n_slots * 8, np.float32) if P["w"] is not None else None
    n_h = int(sum(32 * s[2] for s in subs if s[3] == 16))
    n_q = int(sum(32 * s[2] for s in subs if s[3] == 32))
    n_s = int(sum(32 * s[2] for s in subs if s[3] == 64))
    idx_h = as_np(ptrs[7], n_h * 4, np.uint16)
    idx_q = as_np(ptrs[8], n_q * 2, np.uint16)
    idx_s = as_np(ptrs[9], n_s, np.uint16)
    sub_block = np.zeros(n_subs, dtype=np.int64)
    for s0, s1, blk, _ in units:
        sub_block[s0:s1] = blk
    out_r, out_c, out_w = [], [], []
    for si, (slot_begin, row_begin, n_groups, code) in enumerate(subs):
        blk = int(sub_block[si])
        steps, width = {16: (1, 4), 32: (1, 2), 64: (1, 1)}.get(int(code), (int(code), 8))
        assert narrow or code <= 8
        rows = seg_row[row_begin:row_begin + 32 * n_groups].reshape(n_groups, 32)
        for q in range(n_groups):
            base = slot_begin + q * 32 * steps
            for j in range(steps):
                s = base + j * 32 + np.arange(32)
                if code == 16:
                    ids = idx_h.reshape(-1, 4)[s].astype(np.int64); pad = W
                elif code == 32:
                    ids = idx_q.reshape(-1, 2)[s].astype(np.int64); pad = W
                elif code == 64:
                    ids = idx_s.reshape(-1, 1)[s].astype(np.int64); pad = W
                elif blk < B:
                    ids = idx16.reshape(-1, 8)[s].astype(np.int64); pad = W
                else:
                    ids = idx32.reshape(-1, 8)[s - n_hot].astype(np.int64); pad = P["nv"]
                if bank_order and blk < B and code <= 8:
                    real = ids < W
                    assert (ids < W + 64).all()             # padding = one of the slice's zero columns
                else:
                    real = ids != pad
                    # padding only behind the real entries of a slot
                    assert (real[:, :-1] >= real[:, 1:]).all()
                assert not real[rows[q] < 0].any()          # unused lanes are all padding
                if stats is not None and blk < B and code <= 8:
                    stats["lds"] = stats.get("lds", 0) + 8
                    stats["wavefronts"] = stats.get("wavefronts", 0) + sum(lds_wavefronts(ids[:, k]) for k in range(8))
                if blk < B:
                    assert (ids[real] < W).all()
                    ids = ids + blk * W
                rr = np.repeat(rows[q][:, None], width, 1)
                out_r.append(rr[real]); out_c.append(ids[real])
                if sw is not None and code <= 8:
                    ww = sw.reshape(-1, 8)[s]
                    assert (ww[~real] == 0).all()
                    out_w.append(ww[real])
    r = np.concatenate(out_r) if out_r else np.zeros(0, np.int64)
    c = np.concatenate(out_c) if out_c else np.zeros(0, np.int64)
    w = np.concatenate(out_w) if out_w else None
    return dict(r=r, c=c, w=w, W=W, B=B, subs=subs, units=units, narrow=narrow)


def check_hot(L, g, P, bank_order=False, stats=None):
    H = hot_pieces(L, g, P, bank_order, stats)
    n_hi, nnz_hi = P["seg"][0], P["nnz_hi"]
    rows = np.repeat(np.arange(n_hi), np.diff(P["off"][:n_hi + 1]))
    cols = P["idx"][:nnz_hi].astype(np.int64)
    assert H["r"].size == nnz_hi
    if P["w"] is None:
        o1, o2 = np.lexsort((H["c"], H["r"])), np.lexsort((cols, rows))
        assert (H["r"][o1] == rows[o2]).all() and (H["c"][o1] == cols[o2]).all()
    else:
        o1, o2 = np.lexsort((H["w"], H["c"], H["r"])), np.lexsort((P["w"][:nnz_hi], cols, rows))
        assert (H["r"][o1] == rows[o2]).all() and (H["c"][o1] == cols[o2]).all() and (H["w"][o1] == P["w"][:nnz_hi][o2]).all()
    return H


@pytest.mark.parametrize("weighted", [False, True])
def test_staging_and_piece_layout(emu, monkeypatch, weighted):
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
    monkeypatch.delenv("CUGRAPH_B200_HOT_NARROW", raising=False)
    src, dst, w = make_edges(120_000, 900_000, seed=3 + weighted, weighted=weighted, id_offset=17)
    g = create_graph(emu, src, dst, w)
    P = primary(emu, g)
    check_csr(P, src, dst, w)
    assert P["seg"][0] > 500                      # there are degree >= 32 rows, and several column blocks
    H = check_hot(emu, g, P)
    assert H["B"] >= 2 and not H["narrow"]
    emu.cugraph_graph_free(g)


def test_piece_layout_with_cold_block_and_small_units(emu, monkeypatch):
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
    monkeypatch.setenv("CUGRAPH_B200_HOT_BLOCKS", "1")          # one hot block, the rest of the columns cold (32-bit ids)
    monkeypatch.setenv("CUGRAPH_B200_HOT_UNIT_SLOTS", "1024")
    src, dst, w = make_edges(120_000, 600_000, seed=11)
    g = create_graph(emu, src, dst, w)
    P = primary(emu, g)
    H = check_hot(emu, g, P)
    assert H["B"] == 1 and (H["units"][:, 2] == 1).any()        # cold units exist
    emu.cugraph_graph_free(g)


def test_narrow_piece_layout(emu, monkeypatch):
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
    monkeypatch.setenv("CUGRAPH_B200_HOT_NARROW", "1")
    src, dst, w = make_edges(160_000, 700_000, seed=5)
    g = create_graph(emu, src, dst, w)
    P = primary(emu, g)
    H = check_hot(emu, g, P)
    assert H["narrow"] and all((H["subs"][:, 3] == code).any() for code in (16, 32, 64))
    emu.cugraph_graph_free(g)


@pytest.mark.parametrize("weighted", [False, True])
def test_bank_ordered_piece_layout(emu, monkeypatch, weighted):
    """CUGRAPH_B200_HOT_BANK_ORDER=1: same (row, source[, weight]) multiset, and fewer shared-memory wavefronts per gather"""
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "0")
    src, dst, w = make_edges(120_000, 900_000, seed=21 + weighted, weighted=weighted, id_offset=3)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CUGRAPH_B200_HOT_BANK_ORDER", mode)
        g = create_graph(emu, src, dst, w)
        P = primary(emu, g)
        st = {}
        check_hot(emu, g, P, bank_order=(mode == "1"), stats=st)
        res[mode] = st["wavefronts"] / st["lds"]
        emu.cugraph_graph_free(g)
    print(f"wavefronts per LDS: default order {res['0']:.3f}, bank order {res['1']:.3f}")
    assert res["1"] < 0.6 * res["0"], res


@pytest.mark.parametrize("weighted", [False, True])
def test_low_ell_sweep(emu, monkeypatch, weighted):
    monkeypatch.setenv("CUGRAPH_B200_LOW_ELL", "1")
    src, dst, w = make_edges(60_000, 400_000, seed=21 + weighted, weighted=weighted)
    g = create_graph(emu, src, dst, w)
    P = primary(emu, g)
    nv, n_hi = P["nv"], P["seg"][0]
    r = np.random.default_rng(1)
    x = r.random(nv).astype(np.float32)
    y = np.full(nv, -7.0, dtype=np.float32)
    alpha, init = 0.85, 0.125
    rc = emu.emu_low_ell_sweep(C.c_void_p(emu.handle), g, x.ctypes.data, y.ctypes.data, alpha, init)
    assert rc == 0
    deg = np.diff(P["off"])
    rows = np.repeat(np.arange(P["n_rows"]), deg)
    vals = x[P["idx"]].astype(np.float64) * (P["w"].astype(np.float64) if weighted else 1.0)
    exp = np.bincount(rows, weights=vals, minlength=P["n_rows"]) * alpha + init
    assert (y[:n_hi] == -7.0).all()                              # the degree >= 32 rows belong to the other kernel
    np.testing.assert_allclose(y[n_hi:], exp[n_hi:].astype(np.float32), rtol=2e-6, atol=0)
    assert (deg[n_hi:] < 32).all() and len(set(deg[n_hi:].tolist())) > 10
    emu.cugraph_graph_free(g)


def test_staging_options(emu, monkeypatch):
    """self-loop / multi-edge removal and symmetrisation against numpy"""
    monkeypatch.setenv("CUGRAPH_B200_SWEEP_MIN_EDGES", "1000000000")
    src, dst, _ = make_edges(3_000, 40_000, seed=9)
    g = create_graph(emu, src, dst, None, drop_self_loops=1, drop_multi_edges=1)
    P = primary(emu, g)
    keep = src != dst
    pairs = np.unique(np.stack([dst[keep], src[keep]], 1), axis=0)
    check_csr(P, pairs[:, 1].astype(np.int32), pairs[:, 0].astype(np.int32), None)
    emu.cugraph_graph_free(g)
